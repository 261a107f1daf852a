"""Deterministic synthetic inputs shaped like the reference's workloads.

No datasets exist in this environment, so benchmarks and parity tests use
 * a street-like cloud of 2-D Gaussian surfels (ground plane, facades, clutter boxes, optional
   rigidly moving actors) with the parameter layout `raytracing()` hands to the tracer
   (/root/reference/lib/gaussian_renderer/__init__.py:76-134: means (P,3), scales (P,2) post-exp,
   rotations (P,4) w-first, opacity (P,1) post-sigmoid, shs (P,16,3)), and
 * LiDAR ray grids generated with the reference's range-image convention
   (/root/reference/lib/scene/lidar_sensor.py:395-434): Waymo top LiDAR 64 x 2650, KITTI-360 66 x 1030.

Everything is numpy on the host; callers move it to the device.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

SH_C0 = 0.28209479177387814

WAYMO_H, WAYMO_W = 64, 2650
KITTI_H, KITTI_W = 66, 1030


def waymo_inclinations(H: int = WAYMO_H) -> np.ndarray:
    """Ascending, non-uniform beam inclinations in [-17.6 deg, +2.4 deg] (denser near the horizon)."""
    u = np.linspace(0.0, 1.0, H)
    lo, hi = math.radians(-17.6), math.radians(2.4)
    return (lo + (hi - lo) * (1.0 - (1.0 - u) ** 1.6)).astype(np.float32)


def kitti_inclinations(H: int = KITTI_H) -> np.ndarray:
    return np.linspace(math.radians(-24.9), math.radians(2.0), H).astype(np.float32)


def sensor_pose(frame: int, speed: float = 1.0, yaw_rate: float = 0.002) -> np.ndarray:
    """sensor2world (4,4): straight 10 Hz trajectory along +x (1 m/frame) with a small yaw drift."""
    yaw = yaw_rate * frame
    c, s = math.cos(yaw), math.sin(yaw)
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], np.float32)
    T[:3, 3] = np.array([speed * frame, 0.05 * math.sin(0.1 * frame), 0.0], np.float32)
    return T


def lidar_rays(H: int, W: int, inclinations: np.ndarray, sensor2world: np.ndarray,
               pixel_offset: float = 0.5, angle_offset: float = 0.0):
    """Range-image ray grid: row 0 = highest beam, azimuth ((W-j)-offset)/W*2pi - pi.

    Returns (ray_o (1,3) shared origin, ray_d (H,W,3) unit, world frame)."""
    inc = np.asarray(inclinations, np.float32)[::-1].reshape(H, 1)
    x = (np.arange(W, 0, -1, dtype=np.float32) - np.float32(pixel_offset)) / np.float32(W)
    az = (x * np.float32(2 * math.pi) - np.float32(math.pi) - np.float32(angle_offset)).reshape(1, W)
    d = np.stack([np.cos(inc) * np.cos(az), np.cos(inc) * np.sin(az), np.sin(inc) * np.ones_like(az)], -1)
    d = d.astype(np.float32) @ sensor2world[:3, :3].T.astype(np.float32)
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    o = sensor2world[:3, 3].astype(np.float32).reshape(1, 3)
    return o, np.ascontiguousarray(d.astype(np.float32))


def _quat_from_normal(n: np.ndarray, angle: np.ndarray) -> np.ndarray:
    """Unit quaternion (w,x,y,z) whose rotation maps +z to n, with a random in-plane spin."""
    n = n / np.linalg.norm(n, axis=1, keepdims=True)
    z = np.array([0.0, 0.0, 1.0])
    v = np.cross(np.broadcast_to(z, n.shape), n)
    w = 1.0 + n[:, 2]
    flip = w < 1e-6
    q = np.concatenate([w[:, None], v], 1)
    q[flip] = np.array([0.0, 1.0, 0.0, 0.0])
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    # spin about local z first: q_total = q * q_spin
    cs, sn = np.cos(0.5 * angle), np.sin(0.5 * angle)
    a = q
    bw, bz = cs, sn
    out = np.stack([a[:, 0] * bw - a[:, 3] * bz,
                    a[:, 1] * bw + a[:, 2] * bz,
                    a[:, 2] * bw - a[:, 1] * bz,
                    a[:, 3] * bw + a[:, 0] * bz], 1)
    return out


@dataclass
class Scene:
    means: np.ndarray      # (P,3) f32
    scales: np.ndarray     # (P,2) f32, post-exp
    rots: np.ndarray       # (P,4) f32, w-first (not necessarily unit)
    opac: np.ndarray       # (P,1) f32 in (0,1)
    shs: np.ndarray        # (P,16,3) f32
    actor_id: np.ndarray   # (P,) int32, -1 = static background
    sh_degree: int = 3

    @property
    def P(self) -> int:
        return self.means.shape[0]


def make_street_scene(P: int, seed: int = 0, extent: float = 80.0, n_actors: int = 0,
                      per_actor: int = 10000, sh_degree: int = 3, scale_mult: float = 1.0) -> Scene:
    """Street-like surfel cloud: 60 % ground, 30 % facades, 10 % clutter boxes (+ optional actors)."""
    rng = np.random.default_rng(seed)
    n_act = n_actors * per_actor
    Pb = P - n_act
    assert Pb > 0
    n_ground = int(0.6 * Pb)
    n_fac = int(0.3 * Pb)
    n_clut = Pb - n_ground - n_fac
    pts, nrm, aid = [], [], []

    # ground z ~ -2
    g = np.stack([rng.uniform(-extent, extent + 60.0, n_ground), rng.uniform(-25, 25, n_ground),
                  -2.0 + 0.03 * rng.standard_normal(n_ground)], 1)
    pts.append(g); nrm.append(np.tile([0.0, 0.0, 1.0], (n_ground, 1))); aid.append(np.full(n_ground, -1))

    # facades: buildings on both sides, frontage at |y| in [8, 25]
    nb = 48
    bx0 = rng.uniform(-extent, extent + 40.0, nb); blen = rng.uniform(10, 30, nb)
    by = rng.uniform(8, 25, nb) * np.where(np.arange(nb) % 2 == 0, 1.0, -1.0); bh = rng.uniform(3, 15, nb)
    b = rng.integers(0, nb, n_fac)
    f = np.stack([bx0[b] + blen[b] * rng.uniform(0, 1, n_fac), by[b] + 0.03 * rng.standard_normal(n_fac),
                  -2.0 + (bh[b] + 2.0) * rng.uniform(0, 1, n_fac)], 1)
    fn = np.stack([np.zeros(n_fac), -np.sign(by[b]), np.zeros(n_fac)], 1)
    pts.append(f); nrm.append(fn); aid.append(np.full(n_fac, -1))

    def box_surface(n, centre, size, yaw):
        """n points on the faces of an oriented box, with outward normals."""
        face = rng.integers(0, 6, n); ax = face // 2; sg = np.where(face % 2 == 0, 1.0, -1.0)
        loc = rng.uniform(-0.5, 0.5, (n, 3)) * size
        loc[np.arange(n), ax] = 0.5 * sg * size[ax]
        nl = np.zeros((n, 3)); nl[np.arange(n), ax] = sg
        c, s = math.cos(yaw), math.sin(yaw)
        Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        return loc @ Rz.T + centre, nl @ Rz.T, loc, nl

    # clutter boxes (parked cars, poles ...) within |y| < 8
    ncb = 64
    per = np.full(ncb, n_clut // ncb); per[: n_clut - per.sum()] += 1
    for i in range(ncb):
        size = np.array([rng.uniform(0.5, 4.5), rng.uniform(0.5, 2.0), rng.uniform(0.8, 2.5)])
        centre = np.array([rng.uniform(-extent, extent + 40.0), rng.uniform(-8, 8), -2.0 + 0.5 * size[2]])
        if abs(centre[1]) < 2.5:
            centre[1] = 2.5 * np.sign(centre[1] + 1e-3) + centre[1]
        p_, n_, _, _ = box_surface(int(per[i]), centre, size, rng.uniform(0, math.pi))
        pts.append(p_); nrm.append(n_); aid.append(np.full(int(per[i]), -1))

    # actors: 4.5 x 2 x 1.6 m boxes in their frame-0 pose (see actor_pose for later frames)
    for a in range(n_actors):
        size = np.array([4.5, 2.0, 1.6])
        centre = actor_centre(a, 0)
        p_, n_, _, _ = box_surface(per_actor, centre, size, 0.0)
        pts.append(p_); nrm.append(n_); aid.append(np.full(per_actor, a))

    pts = np.concatenate(pts, 0); nrm = np.concatenate(nrm, 0); aid = np.concatenate(aid, 0).astype(np.int32)
    Pn = pts.shape[0]
    nrm = nrm + 0.1 * rng.standard_normal((Pn, 3))
    rots = _quat_from_normal(nrm, rng.uniform(0, 2 * math.pi, Pn))
    rots = rots * rng.uniform(0.8, 1.25, (Pn, 1))          # the tracer must normalise (auxiliary.h:306)
    s0 = 0.08 * math.sqrt(1.0e6 / max(Pb, 1)) * scale_mult
    scales = np.exp(math.log(s0) + 0.5 * rng.standard_normal((Pn, 2)))
    opac = 1.0 / (1.0 + np.exp(-2.0 * rng.standard_normal((Pn, 1))))
    opac = np.clip(opac, 0.01, 0.999)
    shs = 0.05 * rng.standard_normal((Pn, 16, 3))
    shs[:, 0, 0] = (rng.uniform(0, 1, Pn) - 0.5) / SH_C0
    shs[:, 0, 1] = (1.0 - 0.5) / SH_C0
    shs[:, 0, 2] = (0.0 - 0.5) / SH_C0
    perm = rng.permutation(Pn)                               # no spatial order in the caller's arrays
    f32 = lambda a: np.ascontiguousarray(a[perm].astype(np.float32))
    return Scene(f32(pts), f32(scales), f32(rots), f32(opac), f32(shs), np.ascontiguousarray(aid[perm]), sh_degree)


def actor_centre(a: int, frame: int) -> np.ndarray:
    lane = -1.0 if a % 2 else 1.0
    x0 = -60.0 + 7.5 * a
    return np.array([x0 + lane * 0.8 * frame, 3.2 * lane, -2.0 + 0.8])


def actor_transform(a: int, frame: int):
    """Rigid motion (R (3,3), t (3,)) taking actor a's frame-0 Gaussians to `frame`."""
    t = actor_centre(a, frame) - actor_centre(a, 0)
    return np.eye(3, dtype=np.float32), t.astype(np.float32)


def scene_at_frame(scene: Scene, frame: int) -> Scene:
    """World-frame parameters at `frame` (what raytracing() concatenates, gaussian_renderer:76-134)."""
    if frame == 0 or (scene.actor_id < 0).all():
        return scene
    means = scene.means.copy()
    for a in np.unique(scene.actor_id[scene.actor_id >= 0]):
        R, t = actor_transform(int(a), frame)
        m = scene.actor_id == a
        means[m] = means[m] @ R.T + t
    return Scene(means, scene.scales, scene.rots, scene.opac, scene.shs, scene.actor_id, scene.sh_degree)


def ray_patch(H: int, W: int, frame: int = 0, h0: int = 16, w0: int = 0):
    """An (H, W) patch of the Waymo grid (config #1 uses 64 x 64), spread over all azimuths."""
    o, d = lidar_rays(WAYMO_H, WAYMO_W, waymo_inclinations(), sensor_pose(frame))
    hs = (np.arange(H) * max(1, WAYMO_H // H) + (h0 if H < WAYMO_H else 0)) % WAYMO_H
    ws = (w0 + np.arange(W) * max(1, WAYMO_W // W)) % WAYMO_W
    return o, np.ascontiguousarray(d[np.ix_(hs, ws)])

"""Minimal host-side stand-ins for the reference's scene objects, just enough to drive
`lib.gaussian_renderer.raytracing()` the way train.py / eval.py do.

`GaussianAsset` exposes the accessor surface of the reference's GaussianModel
(/root/reference/lib/scene/gaussian_model.py:25-32 activations, :112-148 accessors) over raw
(pre-activation) leaf parameters; `LidarSensor` exposes `get_range_rays(frame)` /
`sensor_center[frame]` like LiDARSensor (/root/reference/lib/scene/lidar_sensor.py:395-434).
Training orchestration, densification and I/O are out of scope.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import synthetic as syn


class GaussianAsset:
    def __init__(self, scene: "syn.Scene", device="cuda", requires_grad: bool = True, actor_poses=None):
        t = lambda a: torch.tensor(np.ascontiguousarray(a), device=device, dtype=torch.float32).requires_grad_(requires_grad)
        self._xyz = t(scene.means)
        self._scaling = t(np.log(scene.scales))                                  # exp activation
        self._rotation = t(scene.rots)                                           # normalize activation
        op = np.clip(scene.opac, 1e-6, 1 - 1e-6)
        self._opacity = t(np.log(op / (1 - op)))                                 # sigmoid activation
        self._features_dc = t(scene.shs[:, :1, :])
        self._features_rest = t(scene.shs[:, 1:, :])
        self.active_sh_degree = scene.sh_degree
        self.max_sh_degree = 3
        self.actor_poses = actor_poses            # optional {frame: (T (3,), quat (1,4))}

    def parameters(self):
        return [self._xyz, self._scaling, self._rotation, self._opacity, self._features_dc, self._features_rest]

    def get_world_xyz(self, frame=0):
        if self.actor_poses is not None and frame in self.actor_poses:
            T, q = self.actor_poses[frame]
            q = F.normalize(q.reshape(1, 4), dim=1)[0]
            w, x, y, z = q.unbind()
            R = torch.stack([torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)]),
                             torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)]),
                             torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)])])
            return self._xyz @ R.T + T
        return self._xyz

    @property
    def get_opacity(self):
        return torch.sigmoid(self._opacity)

    @property
    def get_scaling(self):
        return torch.exp(self._scaling)

    def get_rotation(self, frame=0):
        if self.actor_poses is not None and frame in self.actor_poses:
            obj = self.actor_poses[frame][1].reshape(1, 4)
        else:
            obj = torch.zeros((1, 4), device=self._rotation.device)
        return obj, F.normalize(self._rotation, dim=1)

    @property
    def get_features(self):
        return torch.cat((self._features_dc, self._features_rest), dim=1)


class LidarSensor:
    """Range-image LiDAR with per-frame poses; rays are produced on the device per call."""

    def __init__(self, H=syn.WAYMO_H, W=syn.WAYMO_W, inclinations=None, device="cuda", pixel_offset=0.5):
        self.H, self.W = H, W
        self.inclinations = syn.waymo_inclinations(H) if inclinations is None else inclinations
        self.device = device
        self.pixel_offset = pixel_offset
        self.sensor_center = {}
        self.sensor2world = {}

    def add_frame(self, frame: int, sensor2world: np.ndarray):
        self.sensor2world[frame] = sensor2world
        self.sensor_center[frame] = torch.tensor(sensor2world[:3, 3], device=self.device, dtype=torch.float32)

    def get_range_rays(self, frame):
        """One kernel (lrt_range_rays) instead of the reference's ~15 torch ops; same (rays_o stride-0 view, rays_d)."""
        from . import native
        if getattr(self, "_nctx", None) is None:
            self._nctx = native.Context(self.device)
        return self._nctx.range_rays(self.H, self.W, self.inclinations, self.sensor2world[frame], self.pixel_offset)

    def range2point(self, frame, range_map):
        """World points (H, W, 3) of a range map: one kernel (lrt_range_points) for lidar_sensor.py:325-393."""
        from . import native
        if getattr(self, "_nctx", None) is None:
            self._nctx = native.Context(self.device)
        rm = torch.as_tensor(range_map, dtype=torch.float32, device=self.device)
        if rm.dim() == 3:
            rm = rm[0] if rm.shape[0] == 1 else rm[..., 0]
        return self._nctx.range_points(rm.detach().contiguous(), self.inclinations, self.sensor2world[frame], self.pixel_offset)

    def inverse_projection_with_range(self, frame, range_map, mask):
        """lidar_sensor.py:170-191: the points of the pixels `mask` keeps, (N, 3). As in the reference no gradient reaches
        `range_map` (it re-wraps the points with torch.tensor(), :183)."""
        pts = self.range2point(frame, range_map)
        if mask.dim() == 2:
            pts = pts[mask.bool()]
        else:
            pts = pts * mask
        return pts.view(-1, 3)

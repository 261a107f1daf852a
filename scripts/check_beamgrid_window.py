"""CPU check of the beam-grid window geometry (lrt_beamgrid.cuh: bg_window / bg_cell_*), development aid.
For a seeded scene, verifies that every (ray, surfel) pair accepted by the relaxed candidate test lies inside
the surfel's (azimuth, elevation) cell window, and reports how tight the windows are.
   python scripts/check_beamgrid_window.py [P] [scale_mult] [tilt_radians]"""
import sys, os, math
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lidar-rt_b200"))
from lidar_rt_b200 import synthetic as syn

f32 = np.float32
P = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
H, W = 64, 256
sc = syn.make_street_scene(P, seed=3, scale_mult=float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
o, d = syn.ray_patch(H, W, frame=1)
d = d.reshape(-1, 3); R = d.shape[0]; o = o.reshape(3)
if len(sys.argv) > 3:   # tilt the whole world: the grid must not depend on the sensor being upright
    th = float(sys.argv[3]); c, s = math.cos(th), math.sin(th)
    Rx = np.array([[1, 0, 0], [0, c, -s], [0, s, c]], f32)
    d = (d @ Rx.T).astype(f32); sc.means[:] = (sc.means - o) @ Rx.T + o
    qx = np.array([math.cos(th / 2), math.sin(th / 2), 0, 0])      # rotate the surfel frames too: q' = qx * q
    a, b = qx, sc.rots.astype(np.float64)
    sc.rots[:] = np.stack([a[0]*b[:,0]-a[1]*b[:,1]-a[2]*b[:,2]-a[3]*b[:,3], a[0]*b[:,1]+a[1]*b[:,0]+a[2]*b[:,3]-a[3]*b[:,2],
                           a[0]*b[:,2]-a[1]*b[:,3]+a[2]*b[:,0]+a[3]*b[:,1], a[0]*b[:,3]+a[1]*b[:,2]-a[2]*b[:,1]+a[3]*b[:,0]], 1).astype(f32)

# records (derive_surfel)
q = sc.rots / np.linalg.norm(sc.rots, axis=1, keepdims=True)
w_, x, y, z = q.T
tu = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y + w_ * z), 2 * (x * z - w_ * y)], 1)
tv = np.stack([2 * (x * y - w_ * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w_ * x)], 1)
n = np.stack([2 * (x * z + w_ * y), 2 * (y * z - w_ * x), 1 - 2 * (x * x + y * y)], 1)
Lu = tu / sc.scales[:, :1]; Lv = tv / sc.scales[:, 1:2]
f = np.sqrt(2 * np.log(sc.opac[:, 0] * 255)) + 0.01
mu = sc.means

# grid plan (k_bg_angles / k_bg_plan)
az = np.arctan2(d[:, 1], d[:, 0]); el = np.arctan2(d[:, 2], np.hypot(d[:, 0], d[:, 1]))
el_hi, el_lo = min(el.max() + 1e-4, math.pi / 2), max(el.min() - 1e-4, -math.pi / 2)
span = max(el_hi - el_lo, 1e-3)
delta = max(math.sqrt(2 * math.pi * span / R), 2 * math.pi / 8192)
NA = max(8, min(int(math.ceil(2 * math.pi / delta)), 8192)); NE = max(1, int(math.ceil(span / delta)))
inv_da, inv_de = NA / (2 * math.pi), NE / span
ca = lambda a: np.floor((a + math.pi) * inv_da).astype(np.int64)
ce = lambda e: np.floor((e - el_lo) * inv_de).astype(np.int64)
ray_a = np.clip(ca(az), 0, NA - 1); ray_e = np.clip(ce(el), 0, NE - 1)
print(f"R={R} NA={NA} NE={NE} cells={NA*NE} delta={math.degrees(delta):.3f} deg")

# windows (bg_window)
lim = (f + 1e-3 * (1 + f)) * 1.0001
su = lim / (Lu ** 2).sum(1); sv = lim / (Lv ** 2).sum(1)
A = Lu * su[:, None]; B = Lv * sv[:, None]
c = mu - o; dist = np.linalg.norm(c, axis=1)
pad = 1e-4 + 2e-5 * dist
ez = np.abs(A[:, 2]) + np.abs(B[:, 2]) + pad
exy = np.hypot(A[:, 0], A[:, 1]) + np.hypot(B[:, 0], B[:, 1]) + pad
rad = np.linalg.norm(A, axis=1) + np.linalg.norm(B, axis=1) + pad
rho = np.hypot(c[:, 0], c[:, 1])
inside = dist <= rad + 2e-3
zlo, zhi = c[:, 2] - ez, c[:, 2] + ez; hlo, hhi = np.maximum(rho - exy, 0), rho + exy
w_el_hi = np.where(inside, math.pi, np.arctan2(zhi, np.where(zhi > 0, hlo, hhi)) + 2e-5)
w_el_lo = np.where(inside, -math.pi, np.arctan2(zlo, np.where(zlo > 0, hhi, hlo)) - 2e-5)
all_az = inside | ~(rho > exy * 1.0001)
valid = np.isfinite(f) & (f >= 0)
e0, e1 = ce(w_el_lo), ce(w_el_hi)
culled = (e1 < 0) | (e0 >= NE) | ~valid
e0c, e1c = np.clip(e0, 0, NE - 1), np.clip(e1, 0, NE - 1)
az0 = np.arctan2(c[:, 1], c[:, 0]); dlt = np.arcsin(np.minimum(exy / np.maximum(rho, 1e-30), 1.0)) + 2e-5
lo = np.zeros(P); hi = np.zeros(P)
for sa in (-1.0, 1.0):
    for sb in (-1.0, 1.0):
        dk = np.arctan2(c[:, 1] + sa * A[:, 1] + sb * B[:, 1], c[:, 0] + sa * A[:, 0] + sb * B[:, 0]) - az0
        dk = np.where(dk > math.pi, dk - 2 * math.pi, np.where(dk < -math.pi, dk + 2 * math.pi, dk))
        lo = np.minimum(lo, dk); hi = np.maximum(hi, dk)
apad = np.arcsin(np.minimum(pad / (np.maximum(rho - exy, 0) + pad), 1.0)) + 2e-5
lo = np.maximum(lo - apad, -dlt); hi = np.minimum(hi + apad, dlt)
a0, a1 = ca(az0 + lo), ca(az0 + hi)
na = np.where(all_az, NA, np.minimum(a1 - a0 + 1, NA)); ia_lo = np.where(all_az, 0, np.mod(a0, NA))
cells = np.where(culled, 0, na * (e1c - e0c + 1))
print(f"surfels: {P}, culled {culled.sum()}, all-azimuth {(all_az & ~culled).sum()}, wide (> 96 az cells) {((na > 96) & ~culled).sum()}, "
      f"mean cells/surfel (kept) {cells[~culled].mean():.1f}, total cells {cells.sum()}")

# brute-force relaxed candidate test (quad_candidate) in chunks
miss = 0; ncand = 0
for g0 in range(0, P, 512):
    g1 = min(P, g0 + 512)
    den = n[g0:g1] @ d.T                               # (G, R)
    num = (n[g0:g1] * c[g0:g1]).sum(1)[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = num / den
        X = t[:, :, None] * d[None] - c[g0:g1, None, :]
        u = (X * Lu[g0:g1, None, :]).sum(2); v = (X * Lv[g0:g1, None, :]).sum(2)
        l = (f[g0:g1] + 1e-3 * (1 + f[g0:g1]))[:, None]
        cand = (t > 0) & (np.abs(u) <= l) & (np.abs(v) <= l) & (t < 1e16)
    gi, ri = np.nonzero(cand); gi += g0
    ncand += len(gi)
    ok_e = (ray_e[ri] >= e0c[gi]) & (ray_e[ri] <= e1c[gi]) & ~culled[gi]
    ok_a = np.mod(ray_a[ri] - ia_lo[gi], NA) < na[gi]
    bad = ~(ok_e & ok_a)
    miss += bad.sum()
    if bad.any() and miss <= 20:
        k = np.nonzero(bad)[0][0]
        print("MISS surfel", gi[k], "ray", ri[k], "ray cell", ray_a[ri[k]], ray_e[ri[k]], "window a", ia_lo[gi[k]], na[gi[k]], "e", e0[gi[k]], e1[gi[k]], "culled", culled[gi[k]])
print(f"candidate pairs {ncand} ({ncand / R:.1f} per ray), missed by the windows: {miss}; tests per candidate ~ {cells.sum() * (R / (NA * NE)) / max(ncand, 1):.1f}")
sys.exit(1 if miss else 0)

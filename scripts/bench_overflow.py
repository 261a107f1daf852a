"""Cost of rays with more candidates than their bin holds (development aid): the 9 000-surfel stack of
tests/test_gpu_parity.py::test_candidates_beyond_the_bin_take_the_overflow_list, 64 rays, forward only.
   bins that hold everything  /  1024-entry bins + overflow list  /  1024-entry bins, per-ray fallback (compositing form 2)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lidar-rt_b200"))
import numpy as np, torch
from lidar_rt_b200 import native

n_stack = int(sys.argv[1]) if len(sys.argv) > 1 else 9000
rng = np.random.default_rng(n_stack)
P = n_stack + 500
means = np.zeros((P, 3), np.float32)
means[:n_stack, 0] = 5.0 + 0.002 * np.arange(n_stack); means[:n_stack, 1:] = 0.01 * rng.standard_normal((n_stack, 2))
means[n_stack:] = rng.uniform(-20, 20, (P - n_stack, 3)) + np.array([30, 0, 0], np.float32)
scales = np.full((P, 2), 0.4, np.float32)
rots = np.tile(np.array([np.cos(np.pi / 4), 0, np.sin(np.pi / 4), 0], np.float32), (P, 1)); rots[n_stack:] = rng.standard_normal((P - n_stack, 4))
opac = np.full((P, 1), 0.012, np.float32)
shs = (0.05 * rng.standard_normal((P, 16, 3))).astype(np.float32); shs[:, 0, :] = 0.5
yy, zz = np.meshgrid(np.linspace(-0.03, 0.03, 8), np.linspace(-0.03, 0.03, 8), indexing="ij")
d = np.stack([np.ones_like(yy), yy, zz], -1).astype(np.float32); d /= np.linalg.norm(d, axis=-1, keepdims=True)
cu = lambda x: torch.as_tensor(np.ascontiguousarray(x), device="cuda")
o, dd, bg = cu(np.zeros((1, 3), np.float32)), cu(d), cu(np.array([0, 0, 1], np.float32))
g = [cu(x) for x in (means, scales, rots, opac, shs)]
ctx = native.Context()
ctx.build(*g[:4])

def timed(label, **opts):
    for k, v in opts.items():
        ctx.set_option(getattr(native, k), v)
    for _ in range(3):
        f = ctx.forward(o, dd, bg, *g, 3, record_hits=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f = ctx.forward(o, dd, bg, *g, 3, record_hits=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"case": label, "rays": 64, "candidates_per_ray": n_stack, "forward_ms": ms, "ms_per_ray": ms / 64}))
    return f["out"].clone()

a = timed("bins of 16384 (everything fits)", OPT_BIN_CAP=16384, OPT_WAVEFRONT_SHADE=3)
b = timed("bins of 1024 + overflow list", OPT_BIN_CAP=1024, OPT_WAVEFRONT_SHADE=3)
c = timed("bins of 1024, per-ray fallback (compositing form 2: no overflow list)", OPT_BIN_CAP=1024, OPT_WAVEFRONT_SHADE=2)
print("outputs identical:", bool(torch.equal(a, b)), bool(torch.equal(a, c)))

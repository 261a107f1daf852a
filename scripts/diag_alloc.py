"""Why does a step sometimes take ~10 ms on the host? Allocator statistics and host/device time per call (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lidar-rt_b200"))
import numpy as np, torch
from lidar_rt_b200 import native, synthetic as syn
dev = torch.device("cuda", 0)
cu = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
BG = cu(np.array([0, 0, 1], np.float32))
which = sys.argv[1] if len(sys.argv) > 1 else "waymo"
if which == "waymo":
    H, W, P = 64, 2650, 2_000_000
    sc = syn.make_street_scene(P, seed=1); inc = syn.waymo_inclinations(); off = 0.5
else:
    H, W = syn.KITTI_H, 1030
    sc = syn.make_street_scene(700_000, seed=3, n_actors=20, per_actor=10_000); inc = syn.kitti_inclinations(H); off = 0.0
g = tuple(map(cu, (sc.means, sc.scales, sc.rots, sc.opac, sc.shs)))
rng = np.random.default_rng(0)
frames = []
for f in range(20):
    o, d = syn.lidar_rays(H, W, inc, syn.sensor_pose(f), pixel_offset=off)
    dL = np.zeros((H, W, 9), np.float32); dL[..., :4] = rng.standard_normal((H, W, 4))
    frames.append((cu(o), cu(d), cu(dL)))
ctx = native.Context(dev)
def stats():
    s = torch.cuda.memory_stats(dev)
    return {k: s.get(k, 0) for k in ("num_device_alloc", "num_device_free", "num_alloc_retries", "reserved_bytes.all.current", "allocated_bytes.all.current")}
outs = []
for f, (ro, rd, gl) in enumerate(frames):
    s0 = stats(); torch.cuda.synchronize(); t = [time.perf_counter()]
    ctx.build(*g[:4]); t.append(time.perf_counter())
    r = ctx.forward(ro, rd, BG, *g, 3); t.append(time.perf_counter())
    ctx.backward(ro, rd, BG, *g, 3, r["out"], gl, hits=r); t.append(time.perf_counter())
    torch.cuda.synchronize(); t.append(time.perf_counter())
    outs.append(r["out"])
    s1 = stats()
    print(which, "frame", f, "host ms build/fwd/bwd/sync", [round(1e3 * (t[i + 1] - t[i]), 2) for i in range(4)],
          "cudaMalloc", s1["num_device_alloc"] - s0["num_device_alloc"], "cudaFree", s1["num_device_free"] - s0["num_device_free"],
          "reserved GB", round(s1["reserved_bytes.all.current"] / 1e9, 2), "allocated GB", round(s1["allocated_bytes.all.current"] / 1e9, 2), flush=True)

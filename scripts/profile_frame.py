"""One or two frames of the hot path (build + forward + backward) for ncu captures and traversal statistics.
   python scripts/profile_frame.py [--gaussians 2000000] [--frames 2] [--stats]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lidar-rt_b200"))
import numpy as np, torch
from lidar_rt_b200 import native, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--gaussians", type=int, default=2_000_000)
ap.add_argument("--frames", type=int, default=2)
ap.add_argument("--stats", action="store_true")
ap.add_argument("--no-backward", action="store_true")
a = ap.parse_args()
BG = np.array([0, 0, 1], np.float32)
cu = lambda x: torch.as_tensor(np.ascontiguousarray(x), device="cuda")
sc = syn.make_street_scene(a.gaussians, seed=1)
means, scales, rots, opac, shs = map(cu, (sc.means, sc.scales, sc.rots, sc.opac, sc.shs))
ctx = native.Context()
inc = syn.waymo_inclinations()
rng = np.random.default_rng(0)
for f in range(a.frames):
    o, d = syn.lidar_rays(64, 2650, inc, syn.sensor_pose(f))
    dL = np.zeros((64, 2650, 9), np.float32); dL[..., :4] = rng.standard_normal((64, 2650, 4))
    ro, rd, g = cu(o), cu(d), cu(dL)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.build(means, scales, rots, opac)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    res = ctx.forward(ro, rd, cu(BG), means, scales, rots, opac, shs, 3, want_slots=True)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    if not a.no_backward:
        ctx.backward(ro, rd, cu(BG), means, scales, rots, opac, shs, 3, res["out"], g, hits=res)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    sl = res["slot_cnt"].cpu().numpy().astype(np.int64); hc = res["hit_cnt"].cpu().numpy()
    print(f"frame {f}: build {1e3*(t1-t0):.2f} ms fwd {1e3*(t2-t1):.2f} ms bwd {1e3*(t3-t2):.2f} ms | slots/ray {np.mean(sl & 0xffff):.1f} "
          f"contrib/ray {hc.mean():.1f} max {hc.max()} overflow {(hc > 64).mean():.3f}")
    if a.stats:
        nodes = sl >> 16
        print(f"   node visits/ray: mean {nodes.mean():.0f} median {np.median(nodes):.0f} p90 {np.percentile(nodes, 90):.0f} max {nodes.max()}")
        rows = nodes.reshape(64, 2650).mean(1)
        print("   node visits by beam row:", np.round(rows[::4]).astype(int))

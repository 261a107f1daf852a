"""Per-kernel times and work counters of one KITTI-shaped dynamic frame (BASELINE config #3); development aid."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lidar-rt_b200"))
import numpy as np, torch
from lidar_rt_b200 import native, synthetic as syn
dev = torch.device("cuda", 0)
cu = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
H, W = syn.KITTI_H, 1030
sc0 = syn.make_street_scene(500_000 + 20 * 10_000, seed=3, n_actors=20, per_actor=10_000)
ctx = native.Context(dev)
ctx.set_option(native.OPT_KERNEL_TIMING, 1)
BG = cu(np.array([0, 0, 1], np.float32))
for variant in [v for v in sys.argv[1:] if v in ("kitti", "kitti_waymo_inc", "refit")]:
    if variant == "kitti":
        inc, off = syn.kitti_inclinations(H), 0.0
    elif variant == "kitti_waymo_inc":      # KITTI grid size with the Waymo elevation band
        inc, off = syn.waymo_inclinations(H), 0.5
    elif variant == "refit":
        inc, off = syn.kitti_inclinations(H), 0.0
    for f in (0, 1, 2, 3) if variant == "refit" else (1, 2):
        sc = syn.scene_at_frame(sc0, f)
        g = tuple(map(cu, (sc.means, sc.scales, sc.rots, sc.opac, sc.shs)))
        o, d = syn.lidar_rays(H, W, inc, syn.sensor_pose(f), pixel_offset=off)
        ctx.build(*g[:4], refit=(variant == "refit" and f > 0)); ctx.kernel_times()
        r = ctx.forward(cu(o), cu(d), BG, *g, 3, want_slots=True)
        torch.cuda.synchronize()
        kt = {k: round(v[0], 3) for k, v in ctx.kernel_times().items()}
        c = (ctypes.c_int * 16)(); ctx.lib.lrt_debug_counters(ctx._h, c)
        hc = r["hit_cnt"].cpu().numpy(); sl = r["slot_cnt"].cpu().numpy()
        print(variant, "frame", f, json.dumps(kt), "counters", list(c), "hits/ray", hc.mean(), "max", hc.max(), "slots/ray", sl.mean(), "max", sl.max(), flush=True)
if "loop" in sys.argv:
    import time
    ctx.set_option(native.OPT_KERNEL_TIMING, 0)
    inc = syn.kitti_inclinations(H)
    rng = np.random.default_rng(1)
    g0 = tuple(map(cu, (sc0.means, sc0.scales, sc0.rots, sc0.opac, sc0.shs)))
    for f in range(8):
        sc = syn.scene_at_frame(sc0, f)
        means = cu(sc.means)
        o, d = syn.lidar_rays(H, W, inc, syn.sensor_pose(f), pixel_offset=0.0)
        dL = np.zeros((H, W, 9), np.float32); dL[..., :4] = rng.standard_normal((H, W, 4))
        ro, rd, gl = cu(o), cu(d), cu(dL)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        t = [time.perf_counter()]
        e[0].record(); ctx.build(means, *g0[1:4], refit=f % 4 != 0); t.append(time.perf_counter())
        e[1].record(); r = ctx.forward(ro, rd, BG, means, *g0[1:], 3); t.append(time.perf_counter())
        e[2].record(); gr = ctx.backward(ro, rd, BG, means, *g0[1:], 3, r["out"], gl, hits=r); t.append(time.perf_counter())
        e[3].record(); torch.cuda.synchronize(); t.append(time.perf_counter())
        print("loop frame", f, "events ms", [round(e[i].elapsed_time(e[i + 1]), 3) for i in range(3)], "host ms", [round(1e3 * (t[i + 1] - t[i]), 3) for i in range(4)], flush=True)

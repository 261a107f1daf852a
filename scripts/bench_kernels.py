"""Per-kernel device times of the default path for a list of option settings (development aid).
   python scripts/bench_kernels.py "" "9=50" "1=3"      # each argument: LIDAR_RT_B200_OPTIONS-style id=value list"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lidar-rt_b200"))
import numpy as np, torch
from lidar_rt_b200 import native, synthetic as syn

P = int(os.environ.get("P", 2_000_000)); K = 12
cu = lambda x: torch.as_tensor(np.ascontiguousarray(x), device="cuda")
sc = syn.make_street_scene(P, seed=1)
means, scales, rots, opac, shs = map(cu, (sc.means, sc.scales, sc.rots, sc.opac, sc.shs))
bg = cu(np.array([0, 0, 1], np.float32))
inc = syn.waymo_inclinations(); rng = np.random.default_rng(0)
frames = []
for f in range(K):
    o, d = syn.lidar_rays(64, 2650, inc, syn.sensor_pose(f))
    dL = np.zeros((64, 2650, 9), np.float32); dL[..., :4] = rng.standard_normal((64, 2650, 4))
    frames.append((cu(o), cu(d), cu(dL)))
ctx = native.Context()
for spec in (sys.argv[1:] or [""]):
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("="); ctx.set_option(int(k), int(v))
    def step(i):
        o, d, dL = frames[i]
        ctx.build(means, scales, rots, opac)
        f = ctx.forward(o, d, bg, means, scales, rots, opac, shs, 3)
        ctx.backward(o, d, bg, means, scales, rots, opac, shs, 3, f["out"], dL, hits=f)
    for i in range(3): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K): step(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    ctx.set_option(native.OPT_KERNEL_TIMING, 1); ctx.kernel_times()
    for i in range(K): step(i)
    kt = ctx.kernel_times(); ctx.set_option(native.OPT_KERNEL_TIMING, 0)
    import ctypes
    c = (ctypes.c_int * 16)(); ctx.lib.lrt_debug_counters(ctx._h, c)
    print(f"[{spec or 'default'}] step {ms:.3f} ms = {64 * 2650 / ms / 1e3:.1f} Mrays/s | " +
          " ".join(f"{k}={v[0] / K:.3f}" for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])) + f" | fallback {c[8]} heavy items {c[9]}", flush=True)

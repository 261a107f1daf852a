"""Frames of the hot path (build + forward + backward) for ncu captures, A/B of tuning options and
traversal statistics (development aid).
   python scripts/profile_frame.py [--gaussians 2000000] [--frames 3] [--ab] [--stats]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lidar-rt_b200"))
import numpy as np, torch
from lidar_rt_b200 import native, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--gaussians", type=int, default=2_000_000)
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--stats", action="store_true")
ap.add_argument("--ab", action="store_true", help="time every combination of the tuning options")
ap.add_argument("--fwd-kernel", type=int, default=4)
ap.add_argument("--flat", action="store_true", help="pass rays as (R,3): no 4x8 tiles")
ap.add_argument("--no-vec", action="store_true")
ap.add_argument("--cap", type=int, default=native.DEFAULT_HIT_CAP)
ap.add_argument("--morton", type=int, default=32)
ap.add_argument("--bwd-kernel", type=int, default=2)
ap.add_argument("--shade", type=int, default=3)
a = ap.parse_args()
BG = np.array([0, 0, 1], np.float32)
cu = lambda x: torch.as_tensor(np.ascontiguousarray(x), device="cuda")
sc = syn.make_street_scene(a.gaussians, seed=1)
means, scales, rots, opac, shs = map(cu, (sc.means, sc.scales, sc.rots, sc.opac, sc.shs))
ctx = native.Context()
inc = syn.waymo_inclinations()
rng = np.random.default_rng(0)
frames = []
for f in range(a.frames):
    o, d = syn.lidar_rays(64, 2650, inc, syn.sensor_pose(f))
    dL = np.zeros((64, 2650, 9), np.float32); dL[..., :4] = rng.standard_normal((64, 2650, 4))
    frames.append((cu(o), cu(d), cu(dL)))
bg = cu(BG)


def run(fwd_kernel, flat, vec, cap, label, verbose=True, morton=None):
    ctx.set_option(native.OPT_MORTON_BITS, morton or a.morton)
    ctx.set_option(native.OPT_FORWARD_KERNEL, fwd_kernel)
    ctx.set_option(native.OPT_BACKWARD_KERNEL, a.bwd_kernel)
    ctx.set_option(native.OPT_WAVEFRONT_SHADE, a.shade)
    ctx.set_option(native.OPT_VECTOR_ATOMICS, int(vec))
    tb, tf, tw = [], [], []
    for f, (ro, rd, g) in enumerate(frames):
        if flat:
            rd = rd.reshape(-1, 3); g = g.reshape(-1, 9)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(); ctx.build(means, scales, rots, opac)
        e[1].record(); res = ctx.forward(ro, rd, bg, means, scales, rots, opac, shs, 3, cap=cap, want_slots=True)
        e[2].record(); ctx.backward(ro, rd, bg, means, scales, rots, opac, shs, 3, res["out"], g, hits=res)
        e[3].record(); torch.cuda.synchronize()
        if f > 0 or a.frames == 1:
            tb.append(e[0].elapsed_time(e[1])); tf.append(e[1].elapsed_time(e[2])); tw.append(e[2].elapsed_time(e[3]))
    import ctypes as _ct
    _c = (_ct.c_int * 16)(); ctx.lib.lrt_debug_counters(ctx._h, _c)
    sl = res["slot_cnt"].cpu().numpy().astype(np.int64); hc = res["hit_cnt"].cpu().numpy()
    print(f"{label:34s} build {np.mean(tb):6.2f} ms  fwd {np.mean(tf):6.2f} ms  bwd {np.mean(tw):6.2f} ms  | slots/ray {np.mean(sl & 0xffff):.1f} "
          f"contrib/ray {hc.mean():.1f} max {hc.max()} overflow {(hc > cap).mean():.4f} fallback rays (last frame) {_c[8]} heavy items {_c[9]}", flush=True)
    if a.stats:
        import ctypes
        st = (ctypes.c_ulonglong * 16)(); ctx.lib.lrt_debug_stats(st, 1); st = np.array(list(st), np.float64) / (a.frames * hc.shape[0])
        print(f"   per ray: evals by level {np.round(st[:8], 1)} quad tests {st[8]:.1f} quad hits {st[9]:.1f} re-evals {st[10]:.1f} rounds {st[11]:.2f}")
        nodes = sl >> 16
        print(f"   node evaluations/ray: mean {nodes.mean():.0f} median {np.median(nodes):.0f} p90 {np.percentile(nodes, 90):.0f} max {nodes.max()}")
        print("   by beam row (every 4th):", np.round(nodes.reshape(64, 2650).mean(1)[::4]).astype(int))
    return res


if a.ab:
    ref = None
    for fk in (3, 0, 1, 2):
        for flat in ((True, False) if fk in (0, 3) else (False,)):
            res = run(fk, flat, True, 128, f"fwd_kernel={fk} tiles={'no' if flat else '4x8'} vec=1 cap=128")
            out = res["out"].reshape(-1, 9)
            if ref is None:
                ref = out.clone()
            else:
                print("      identical to first config:", bool(torch.equal(ref, out)))
    run(2, False, True, 128, "fwd_kernel=2 tiles=4x8 morton=30", morton=30)
    run(2, True, True, 128, "fwd_kernel=2 tiles=no")
else:
    run(a.fwd_kernel, a.flat, not a.no_vec, a.cap, f"fwd_kernel={a.fwd_kernel} tiles={'no' if a.flat else '4x8'} vec={int(not a.no_vec)} cap={a.cap} morton={a.morton} bwd_kernel={a.bwd_kernel} shade={a.shade}")

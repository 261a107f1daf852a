"""GPU diagnostic for the wavefront forward (development aid)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lidar-rt_b200"))
import numpy as np, torch
from lidar_rt_b200 import native, synthetic as syn
BG = np.array([0, 0, 1], np.float32)
cu = lambda x: torch.as_tensor(np.ascontiguousarray(x), device="cuda")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
sc = syn.make_street_scene(P, seed=1)
means, scales, rots, opac, shs = map(cu, (sc.means, sc.scales, sc.rots, sc.opac, sc.shs))
ctx = native.Context()
o, d = syn.lidar_rays(64, 2650, syn.waymo_inclinations(), syn.sensor_pose(0))
ro, rd, bg = cu(o), cu(d), cu(BG)
ctx.build(means, scales, rots, opac)
res = {}
for k in (1, 3):
    ctx.set_option(native.OPT_FORWARD_KERNEL, k)
    f = ctx.forward(ro, rd, bg, means, scales, rots, opac, shs, 3, want_slots=True)
    torch.cuda.synchronize()
    res[k] = {n: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for n, v in f.items()}
    if k == 3:
        cnt = (ctypes.c_int * 16)(); ctx.lib.lrt_debug_counters(ctx._h, cnt); print("wavefront counters (items per level 0..7, [8]=fallback rays):", list(cnt), "R =", 64 * 2650)
a, b = res[1], res[3]
oa, ob = a["out"].reshape(-1, 9), b["out"].reshape(-1, 9)
bad = np.where((oa != ob).any(1))[0]
print("rays with differing out:", len(bad), "of", oa.shape[0], "| hit_cnt differs:", int((a["hit_cnt"] != b["hit_cnt"]).sum()), "| slot_cnt differs:", int((a["slot_cnt"] != b["slot_cnt"]).sum()))
for r in bad[:8]:
    ca, cb = a["hit_cnt"][r], b["hit_cnt"][r]
    la, lb = list(a["hit_gidx"][:min(ca, 128), r]), list(b["hit_gidx"][:min(cb, 128), r])
    first = next((i for i in range(min(len(la), len(lb))) if la[i] != lb[i]), None)
    print(f"ray {r}: k1 cnt={ca} slots={a['slot_cnt'][r]} | k3 cnt={cb} slots={b['slot_cnt'][r]} first differing hit idx {first}")
    if first is not None:
        print("   k1:", la[max(0, first - 2):first + 4], np.round(a["hit_t"][max(0, first - 2):first + 4, r], 5))
        print("   k3:", lb[max(0, first - 2):first + 4], np.round(b["hit_t"][max(0, first - 2):first + 4, r], 5))
    print("   out k1", oa[r, :5], "k3", ob[r, :5])

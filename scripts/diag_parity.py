"""GPU diagnostic: where do CUDA hit lists differ from the oracle? (development aid, uses oracle/)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lidar-rt_b200"))
import numpy as np, torch
from lidar_rt_b200 import native, synthetic as syn
from oracle.oracle import Oracle, ORC_BVH

BG = np.array([0, 0, 1], np.float32)
orc = Oracle(False)
ctx = native.Context()
cu = lambda x: torch.as_tensor(np.ascontiguousarray(x), device="cuda")


def derive(sc, g):
    f32 = np.float32
    q = sc.rots[g].astype(f32); nrm = f32(q[0]*q[0] + q[1]*q[1] + q[2]*q[2] + q[3]*q[3]); inv = f32(1)/np.sqrt(nrm)
    w, x, y, z = (q*inv).astype(f32)
    tu = np.array([1-2*(y*y+z*z), 2*(x*y+w*z), 2*(x*z-w*y)], f32); tv = np.array([2*(x*y-w*z), 1-2*(x*x+z*z), 2*(y*z+w*x)], f32)
    n = np.array([2*(x*z+w*y), 2*(y*z-w*x), 1-2*(x*x+y*y)], f32)
    sx, sy = sc.scales[g]; op = sc.opac[g, 0]
    f = np.sqrt(f32(2)*np.log(op*f32(255))) + f32(0.01)
    return sc.means[g], tu, tv, n, sx, sy, op, f


def explain(sc, g, o, d):
    mu, tu, tv, n, sx, sy, op, f = derive(sc, g)
    o64, d64 = o.astype(np.float64), d.astype(np.float64)
    t = float(n.astype(np.float64) @ (mu.astype(np.float64) - o64) / (n.astype(np.float64) @ d64))
    x = o64 + t*d64; r = x - mu
    u = float(tu @ r / sx); v = float(tv @ r / sy)
    G = np.exp(-0.5*(u*u+v*v)); a = min(0.99, op*G)
    e = np.abs(tu)*sx*f + np.abs(tv)*sy*f
    return f"g={g} t={t:.6f} u={u:.5f} v={v:.5f} f={f:.5f} |u|/f={abs(u)/f:.6f} |v|/f={abs(v)/f:.6f} alpha={a:.6f} (1/255={1/255:.6f}) op={op:.4f} half-extent={e}"


for (P, H, W, seed, D, fr) in [(20000, 32, 64, 2, 3, 0), (200_000, 16, 512, 5, 3, 2)]:
    sc = syn.make_street_scene(P, seed=seed)
    o, d = syn.ray_patch(H, W, frame=fr)
    means, scales, rots, opac, shs = map(cu, (sc.means, sc.scales, sc.rots, sc.opac, sc.shs))
    ctx.build(means, scales, rots, opac)
    f = ctx.forward(cu(o), cu(d), cu(BG), means, scales, rots, opac, shs, D, cap=96, want_slots=True)
    fo = orc.forward(o, d, BG, sc.means, sc.scales, sc.rots, sc.opac, sc.shs, D, flags=ORC_BVH, cap=96)
    cnt = f["hit_cnt"].cpu().numpy(); hg = f["hit_gidx"].cpu().numpy(); ht = f["hit_t"].cpu().numpy(); sl = f["slot_cnt"].cpu().numpy()
    out = f["out"].reshape(-1, 9).cpu().numpy()
    dd = d.reshape(-1, 3)
    bad = [r for r in range(cnt.shape[0]) if cnt[r] != fo["hit_cnt"][r] or sl[r] != fo["slot_cnt"][r]
           or list(hg[:min(cnt[r], 96), r]) != list(fo["hit_list"][r, :min(cnt[r], 96)])]
    print(f"=== P={P} rays={cnt.shape[0]} levels={ctx.info().levels}: {len(bad)} rays differ; max out err {np.abs(out-fo['out']).max():.3e}")
    for r in bad[:6]:
        a = list(hg[:min(cnt[r], 96), r]); b = list(fo["hit_list"][r, :min(fo["hit_cnt"][r], 96)])
        print(f"ray {r}: cuda cnt={cnt[r]} slots={sl[r]} | oracle cnt={fo['hit_cnt'][r]} slots={fo['slot_cnt'][r]} d={dd[r]}")
        print("   cuda  :", a[:40]); print("   oracle:", b[:40])
        print("   cuda t:", np.round(ht[:min(cnt[r], 40), r], 4))
        for g in sorted(set(a) ^ set(b))[:6]:
            print("   only in", "cuda  " if g in a else "oracle", explain(sc, int(g), o[0], dd[r]))
        first = next((i for i in range(min(len(a), len(b))) if a[i] != b[i]), None)
        print("   first differing position:", first, "out cuda", out[r, :5], "oracle", fo["out"][r, :5])

"""What would change if every hit were evaluated once, with its depth taken from the ORIGINAL origin, instead of round by round
from the re-based origin? CPU-only analysis with the C oracle: reference-mode arithmetic against ORC_FLAT on the
known-answer scenes, the small scene, BASELINE config #1 and seeded street scenes, and both against the goldens of the reference
on real OptiX.   python scripts/analyse_flat_depth.py [--big]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lidar-rt_b200"))
import numpy as np
from lidar_rt_b200 import synthetic as syn
from oracle.oracle import ORC_BVH, ORC_FLAT, Oracle

orc = Oracle(False)
BG = np.array([0, 0, 1], np.float32)
rows = []


def compare(name, o, d, sc, D, dL=None):
    a = (o, d, BG, sc.means, sc.scales, sc.rots, sc.opac, sc.shs, D)
    f0 = orc.forward(*a, flags=ORC_BVH, cap=256); f1 = orc.forward(*a, flags=ORC_BVH | ORC_FLAT, cap=256)
    R = f0["out"].shape[0]
    same_cnt = f0["hit_cnt"] == f1["hit_cnt"]
    lists_equal = np.array([same_cnt[r] and np.array_equal(f0["hit_list"][r, :min(f0["hit_cnt"][r], 256)], f1["hit_list"][r, :min(f1["hit_cnt"][r], 256)]) for r in range(R)])
    err = np.abs(f0["out"] - f1["out"]) / (1.0 + np.abs(f0["out"]))
    row = {"case": name, "rays": int(R), "rays_with_other_hit_list": int((~lists_equal).sum()), "rays_with_other_slot_count": int((f0["slot_cnt"] != f1["slot_cnt"]).sum()),
           "max_rel_err_all_rays": float(err.max()), "max_rel_err_same_list_rays": float(err[lists_equal].max()) if lists_equal.any() else None,
           "rays_beyond_1e-6": int((err.max(1) > 1e-6).sum()), "rays_beyond_1e-4": int((err.max(1) > 1e-4).sum()),
           "mean_hits_per_ray": float(f0["hit_cnt"].mean())}
    if dL is not None:
        g0 = orc.backward(*a, f0["out"], dL, flags=ORC_BVH); g1 = orc.backward(*a, f1["out"], dL, flags=ORC_BVH | ORC_FLAT)
        row["grad_rel_l2"] = {k: float(np.linalg.norm(g0[k] - g1[k]) / max(np.linalg.norm(g0[k]), 1e-30)) for k in ("means", "shs", "opac", "scales", "rots")}
    rows.append(row)
    print(json.dumps(row), flush=True)
    return f0, f1


# config #1 and street scenes of growing density
sc = syn.make_street_scene(10_000, seed=0); o, d = syn.ray_patch(64, 64)
compare("config1_10k_4096rays", o, d.reshape(-1, 3), sc, 3)
for P, H, W, seed in ((60_000, 32, 96, 12), (200_000, 16, 512, 5)) + (((1_000_000, 32, 512, 2),) if "--big" in sys.argv else ()):
    sc = syn.make_street_scene(P, seed=seed); o, d = syn.ray_patch(H, W, frame=2)
    rng = np.random.default_rng(seed); dL = np.zeros((H * W, 9), np.float32); dL[:, :4] = rng.standard_normal((H * W, 4))
    compare(f"street_{P // 1000}k_{H * W}rays", o, d.reshape(-1, 3), sc, 3, dL)
# long rays: many rounds per ray (stack of surfels), where re-basing happens most often
n = 3000
rng = np.random.default_rng(3)
means = np.zeros((n, 3), np.float32); means[:, 0] = 5.0 + 0.02 * np.arange(n); means[:, 1:] = 0.01 * rng.standard_normal((n, 2))
stack = syn.Scene(means, np.full((n, 2), 0.4, np.float32), np.tile(np.array([np.cos(np.pi / 4), 0, np.sin(np.pi / 4), 0], np.float32), (n, 1)),
                  np.full((n, 1), 0.012, np.float32), (0.05 * rng.standard_normal((n, 16, 3))).astype(np.float32), np.full(n, -1, np.int32), 3)
yy, zz = np.meshgrid(np.linspace(-0.03, 0.03, 16), np.linspace(-0.03, 0.03, 16), indexing="ij")
dd = np.stack([np.ones_like(yy), yy, zz], -1).astype(np.float32).reshape(-1, 3); dd /= np.linalg.norm(dd, axis=-1, keepdims=True)
compare("stack_3000_surfels_256rays_(~58_rounds_per_ray)", np.zeros((1, 3), np.float32), dd, stack, 3)

# against the reference on real OptiX (goldens): does the flat mode fit them as well as the reference-mode oracle does?
gp = os.path.join(ROOT, "tests", "golden", "optix_b200.npz")
if os.path.exists(gp):
    G = np.load(gp)
    keys = sorted({k.split("/")[0] for k in G.files})
    print("golden groups:", keys[:6], "...", len(keys))
out = os.path.join(ROOT, "profiles", "r1_j_flat_depth_analysis.json")
json.dump(rows, open(out, "w"), indent=1)
print("wrote", out)

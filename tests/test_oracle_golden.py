"""CPU tests: the C restatement (oracle/) against the reference's own code.

Golden vectors in tests/golden/ were produced by oracle/make_golden.py from
 * oracle/_ref  = the reference's forward.cu / backward.cu compiled unmodified as host code, and
 * the reference's Python helpers (eval_sh, build_rotation, build2DRectangle, get_range_rays)
   executed unmodified on the CPU.
"""
import numpy as np
import pytest

from conftest import BG, assert_close, grad_close, load_golden
from lidar_rt_b200 import synthetic as syn
from oracle.oracle import ORC_BVH, ORC_FIX_BG, ORC_TRIANGLES, Oracle, Ref, ref_available

FWD_ATOL, FWD_RTOL = 2e-5, 2e-5      # oracle (ideal-arithmetic restatement) vs reference code
GRAD_REL = 1e-3


def test_sh_matches_reference_eval_sh(oracle32):
    g = load_golden("ref_python.npz")
    sh, dirs = g["sh_coeffs"], g["sh_dirs"]
    for deg in range(4):
        want = g[f"sh_eval_deg{deg}"] + 0.5
        want[:, 0] = np.maximum(want[:, 0], 0.0)            # forward.cu:108-110
        got = np.stack([oracle32.sh_eval(deg, dirs[i], sh[i])[0] for i in range(sh.shape[0])])
        assert_close(got, want, 2e-6, 2e-6, f"SH degree {deg}")


def test_rectangles_match_reference_build2DRectangle(oracle32):
    g = load_golden("ref_python.npz")
    v = oracle32.build_rectangles(g["rect_means"], g["rect_scales"], g["quat"], g["rect_opac"])
    assert_close(v, g["rect_vertices"], 2e-5, 2e-6, "proxy quad corners")
    f = g["rect_faces"].reshape(-1, 2, 3)
    base = 4 * np.arange(f.shape[0])[:, None]
    assert np.array_equal(f[:, 0], base + np.array([0, 1, 2])) and np.array_equal(f[:, 1], base + np.array([2, 3, 1]))


def test_lidar_rays_match_reference_get_range_rays():
    g = load_golden("ref_python.npz")
    for tag, off in (("waymo", 0.5), ("kitti", 0.0)):
        want_o, want_d = g[f"rays_{tag}_o"], g[f"rays_{tag}_d"]
        H, W = want_d.shape[:2]
        inc = g[f"rays_{tag}_inc"]
        if inc.shape[0] == 2:      # two bounds -> linear ramp ((H - h) - off) / H, lidar_sensor.py:412-424
            h = (np.arange(H, 0, -1, dtype=np.float32) - off) / H
            inc = (h * (inc[1] - inc[0]) + inc[0])[::-1]
        o, d = syn.lidar_rays(H, W, inc, g[f"rays_{tag}_pose"], pixel_offset=off)
        assert_close(d, want_d, 2e-6, 0, f"{tag} ray directions")
        assert_close(np.broadcast_to(o, want_o.reshape(-1, 3).shape), want_o.reshape(-1, 3), 0, 0, f"{tag} origins")


def _kat(g, name):
    sc = {k: g[f"{name}/{k}"] for k in ("means", "scales", "rots", "opac", "shs")}
    return sc, g[f"{name}/ray_o"], g[f"{name}/ray_d"], int(g[f"{name}/D"]), g[f"{name}/dL"]


@pytest.mark.parametrize("flags", [0, ORC_TRIANGLES, ORC_BVH], ids=["analytic", "triangles", "bvh"])
def test_known_answer_cases_vs_reference(oracle32, flags):
    g = load_golden("ref_kat.npz")
    for name in g["names"]:
        sc, o, d, D, dL = _kat(g, name)
        args = (o, d, BG, sc["means"], sc["scales"], sc["rots"], sc["opac"], sc["shs"], D)
        f = oracle32.forward(*args, flags=flags)
        assert_close(f["out"], g[f"{name}/out"], FWD_ATOL, FWD_RTOL, f"{name} forward")
        assert_close(f["accum_w"], g[f"{name}/accum_w"], 1e-5, 1e-5, f"{name} accum weights")
        b = oracle32.backward(*args, g[f"{name}/out"], dL, flags=flags)
        for k in ("means", "shs", "opac", "scales", "rots"):
            grad_close(b[k], g[f"{name}/g_{k}"], GRAD_REL, f"{name} d_{k}")


def test_single_gaussian_closed_form(oracle32):
    """alpha = min(.99, o), depth = alpha t, colour = alpha (c0 + .5) + (1 - alpha) bg."""
    g = load_golden("ref_kat.npz")
    sc, o, d, D, _ = _kat(g, "single_onaxis")
    out = oracle32.forward(o, d, BG, sc["means"], sc["scales"], sc["rots"], sc["opac"], sc["shs"], D)["out"][0]
    a = float(sc["opac"][0, 0]); t = 5.0
    c = np.array([0.3, 0.1, -0.2]) + 0.5
    want = np.concatenate([a * c + (1 - a) * BG, [a * t, a, 0, 0, 0, 1 - a]])
    assert_close(out, want, 2e-6, 2e-6, "closed form")
    assert_close(g["single_onaxis/out"][0], want, 2e-6, 2e-6, "reference vs closed form")


def test_round_boundaries_and_epsilon_gap(oracle32):
    g = load_golden("ref_kat.npz")
    # 15 / 16 / 17 / 32 / 33 / 40 stacked surfels: slots consumed = all hits; contributing = those with alpha >= 1/255
    for n in (15, 16, 17, 32, 33, 40):
        sc, o, d, D, _ = _kat(g, f"stack_{n}")
        f = oracle32.forward(o, d, BG, sc["means"], sc["scales"], sc["rots"], sc["opac"], sc["shs"], D)
        assert f["slot_cnt"][0] == n and f["hit_cnt"][0] == n
    sc, o, d, D, _ = _kat(g, "epsilon_gap")
    f = oracle32.forward(o, d, BG, sc["means"], sc["scales"], sc["rots"], sc["opac"], sc["shs"], D)
    # the 17th surfel sits 4e-6 behind the 16th: inside STEP_EPSILON, lost by the re-based round (forward.cu:288-291)
    assert f["hit_cnt"][0] == 18 and 16 not in f["hit_list"][0, :18]
    sc, o, d, D, _ = _kat(g, "terminate")
    f = oracle32.forward(o, d, BG, sc["means"], sc["scales"], sc["rots"], sc["opac"], sc["shs"], D)
    T = f["out"][0, 8]
    assert T >= 1e-4 and T * (1 - 0.95) < 1e-4 * 1.0001      # the hit that would cross 1e-4 is not composited
    sc, o, d, D, _ = _kat(g, "near_cutoff_0p2")
    f = oracle32.forward(o, d, BG, sc["means"], sc["scales"], sc["rots"], sc["opac"], sc["shs"], D)
    assert list(f["hit_list"][0, :2]) == [2, 3] and f["hit_cnt"][0] == 2 and f["slot_cnt"][0] == 4
    sc, o, d, D, _ = _kat(g, "below_alpha_min")
    f = oracle32.forward(o, d, BG, sc["means"], sc["scales"], sc["rots"], sc["opac"], sc["shs"], D)
    assert list(f["hit_list"][0, :2]) == [0, 2] and f["hit_cnt"][0] == 2


@pytest.mark.parametrize("flags", [0, ORC_TRIANGLES], ids=["analytic", "triangles"])
def test_small_scene_forward_backward_vs_reference(oracle32, flags):
    g = load_golden("ref_scene_small.npz")
    args = (g["ray_o"], g["ray_d"], BG, g["means"], g["scales"], g["rots"], g["opac"], g["shs"], int(g["D"]))
    f = oracle32.forward(*args, flags=flags)
    assert_close(f["out"], g["out"], FWD_ATOL, FWD_RTOL, "forward")
    assert_close(f["accum_w"], g["accum_w"], 2e-5, 2e-5, "accum")
    b = oracle32.backward(*args, g["out"], g["dL"], flags=flags)
    for k in ("means", "shs", "opac", "scales", "rots"):
        grad_close(b[k], g[f"g_{k}"], GRAD_REL, f"d_{k}")


def test_config1_forward_vs_reference_and_bvh_filter_is_exact(oracle32):
    """BASELINE config #1: 10k Gaussians, 64x64 rays, forward on the CPU."""
    g = load_golden("ref_cfg1_forward.npz")
    sc = syn.make_street_scene(int(g["P"]), seed=int(g["seed"]))
    o, d = syn.ray_patch(64, 64)
    chk = np.array([float(np.abs(v.astype(np.float64)).sum()) for v in (sc.means, sc.scales, sc.rots, sc.opac, sc.shs, d)])
    assert np.allclose(chk, g["input_checksums"], rtol=1e-9), "synthetic generator drifted from the fixture"
    args = (o, d, BG, sc.means, sc.scales, sc.rots, sc.opac, sc.shs, 3)
    brute = oracle32.forward(*args, flags=0)
    assert_close(brute["out"], g["out"], 2e-4, 2e-5, "config #1 forward")      # depth reaches 80 m
    assert_close(brute["accum_w"], g["accum_w"], 5e-5, 2e-5, "config #1 accum")
    bvh = oracle32.forward(*args, flags=ORC_BVH)
    assert np.array_equal(brute["out"], bvh["out"]) and np.array_equal(brute["hit_list"], bvh["hit_list"])
    assert np.array_equal(brute["slot_cnt"], bvh["slot_cnt"])


def test_bvh_filter_is_exact_with_scale_modifier_and_per_ray_origins(oracle32):
    sc = syn.make_street_scene(8000, seed=31, extent=40.0)
    _, d = syn.ray_patch(8, 32)
    d = d.reshape(-1, 3)
    o = np.random.default_rng(31).uniform(-1, 1, d.shape).astype(np.float32)
    for mod in (0.7, 1.0, 1.4):
        args = (o, d, BG, sc.means, sc.scales, sc.rots, sc.opac, sc.shs, 2)
        a = oracle32.forward(*args, flags=0, scale_modifier=mod, cap=96); b = oracle32.forward(*args, flags=ORC_BVH, scale_modifier=mod, cap=96)
        assert np.array_equal(a["out"], b["out"]) and np.array_equal(a["hit_list"], b["hit_list"]) and np.array_equal(a["slot_cnt"], b["slot_cnt"])


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_random_scenes(oracle32):
    ref = Ref()
    rng = np.random.default_rng(123)
    for seed in (1, 2):
        sc = syn.make_street_scene(3000, seed=seed, extent=30.0, scale_mult=0.6)
        o, d = syn.ray_patch(16, 48, frame=seed)
        args = (o, d, BG, sc.means, sc.scales, sc.rots, sc.opac, sc.shs, seed + 1)
        r = ref.forward(*args)
        f = oracle32.forward(*args, flags=ORC_BVH)
        assert_close(f["out"], r["out"], 1e-4, 2e-5, "forward")
        dL = np.zeros((16 * 48, 9), np.float32); dL[:, :4] = rng.standard_normal((16 * 48, 4))
        gr = ref.backward(*args, r["out"], dL)
        go = oracle32.backward(*args, r["out"], dL, flags=ORC_BVH)
        for k in gr:
            grad_close(go[k], gr[k], GRAD_REL, f"seed {seed} d_{k}")


def _loss64(orc, args, dL):
    out = orc.forward(*args)["out"]
    return float((out * dL).sum())


def test_backward_is_the_gradient_fp64_finite_differences(oracle64):
    """Clean mode == finite differences; reference mode double-counts the background (backward.cu:595-598)."""
    rng = np.random.default_rng(9)
    P = 40
    means = rng.uniform(-0.6, 0.6, (P, 3)); means[:, 2] = rng.uniform(2, 6, P)
    rots = rng.standard_normal((P, 4)); rots /= np.linalg.norm(rots, axis=1, keepdims=True)
    scales = np.exp(rng.normal(-0.3, 0.3, (P, 2))); opac = rng.uniform(0.1, 0.8, (P, 1))
    shs = 0.3 * rng.standard_normal((P, 16, 3)); shs[:, 0, 0] += 1.0     # keep channel 0 unclamped
    d = rng.standard_normal((12, 3)) * np.array([0.08, 0.08, 0]) + np.array([0, 0, 1.0])
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = np.array([[0.02, -0.01, 0.0]])
    bg = np.array([0.2, 0.5, 1.0])
    dL = np.zeros((12, 9)); dL[:, :4] = rng.standard_normal((12, 4))
    params = dict(means=means, scales=scales, rots=rots, opac=opac, shs=shs)

    def args_of(p):
        return (o, d, bg, p["means"], p["scales"], p["rots"], p["opac"], p["shs"], 3)

    out0 = oracle64.forward(*args_of(params))["out"]
    g_clean = oracle64.backward(*args_of(params), out0, dL, flags=ORC_FIX_BG)
    g_ref = oracle64.backward(*args_of(params), out0, dL, flags=0)
    eps = 1e-6
    for key in ("means", "scales", "rots", "opac", "shs"):
        x = params[key]
        idxs = [tuple(rng.integers(0, s) for s in x.shape) for _ in range(25)]
        for idx in idxs:
            p = {k: v.copy() for k, v in params.items()}
            p[key][idx] += eps; up = _loss64(oracle64, args_of(p), dL)
            p[key][idx] -= 2 * eps; dn = _loss64(oracle64, args_of(p), dL)
            fd = (up - dn) / (2 * eps)
            an = g_clean[key].reshape(x.shape)[idx]
            if key == "rots":      # VJP is w.r.t. the normalised quaternion, no gradient through |q| (auxiliary.h:389)
                gi = g_clean["rots"][idx[0]]; qi = rots[idx[0]]
                an = (gi - qi * (qi @ gi))[idx[1]]
            assert abs(fd - an) <= 1e-5 + 1e-4 * abs(fd), f"{key}{idx}: fd {fd} vs analytic {an}"
    # the reference's extra term: -T_final/(1-alpha) * sum_ch dL[ch] bg[ch] per contributing hit
    assert np.abs(g_ref["opac"] - g_clean["opac"]).max() > 1e-3
    dL0 = dL.copy(); dL0[:, :3] = 0
    a = oracle64.backward(*args_of(params), out0, dL0, flags=0); b = oracle64.backward(*args_of(params), out0, dL0, flags=ORC_FIX_BG)
    assert all(np.array_equal(a[k], b[k]) for k in a)


def test_flat_depth_analysis_mode_stays_within_tolerance_of_reference_mode(oracle32):
    """ORC_FLAT (one hit list from the original origin, depth = t) is an analysis mode: on config #1 it must give the
    reference-mode hit lists on every ray and outputs within the 1e-4 parity tolerance."""
    from oracle.oracle import ORC_FLAT
    sc = syn.make_street_scene(10_000, seed=0)
    o, d = syn.ray_patch(64, 64)
    a = (o, d.reshape(-1, 3), BG, sc.means, sc.scales, sc.rots, sc.opac, sc.shs, 3)
    f0 = oracle32.forward(*a, flags=ORC_BVH, cap=128); f1 = oracle32.forward(*a, flags=ORC_BVH | ORC_FLAT, cap=128)
    assert np.array_equal(f0["hit_cnt"], f1["hit_cnt"]) and np.array_equal(f0["hit_list"], f1["hit_list"])
    assert (np.abs(f0["out"] - f1["out"]) / (1.0 + np.abs(f0["out"]))).max() < 1e-4

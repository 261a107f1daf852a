"""GPU parity tests: the CUDA path, called through the C ABI, against
 (a) the golden vectors produced by the reference's own code (tests/golden, oracle/make_golden.py),
 (b) the C oracle on the same seeded inputs (bit-exact hit indices, fp32 tolerance on floats),
 (c) size-independent properties at BASELINE.json's full size.
Tolerances (north_star: fp32 within 1e-4 on depth/intensity, bit-exact hit indices):
   forward floats  |a-b| <= 1e-4 + 1e-4 |b|   vs the reference goldens (different intersector arithmetic)
                   |a-b| <= 2e-5 + 2e-5 |b|   vs the oracle (same arithmetic except expf/logf ulps)
   gradients       max|a-b| / max|b| <= 2e-3  (float atomics reorder sums; reference says the same, train.py:51-64)
"""
import os

import numpy as np
import pytest
import torch

from conftest import BG, assert_close, grad_close, load_golden
from lidar_rt_b200 import synthetic as syn
from oracle.oracle import ORC_BVH, Oracle

pytestmark = pytest.mark.gpu

REF_ATOL, REF_RTOL = 1e-4, 1e-4
ORC_ATOL, ORC_RTOL = 2e-5, 2e-5
GRAD_REL = 2e-3


@pytest.fixture(scope="module")
def ctx():
    from lidar_rt_b200 import native
    c = native.Context()
    yield c
    c.close()


def cu(x):
    return torch.as_tensor(np.ascontiguousarray(x), device="cuda")


def run_cuda(ctx, o, d, sc, D, dL=None, cap=64, use_lists=True, refit=False, flags=0, mod=1.0):
    means, scales, rots, opac, shs = (cu(sc[k]) for k in ("means", "scales", "rots", "opac", "shs"))
    ctx.build(means, scales, rots, opac, scale_modifier=mod, refit=refit)
    ro, rd, bg = cu(o), cu(d), cu(BG)
    f = ctx.forward(ro, rd, bg, means, scales, rots, opac, shs, D, scale_modifier=mod, cap=cap, want_slots=True)
    res = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in f.items()}
    res["out"] = res["out"].reshape(-1, 9)
    if dL is not None:
        g = ctx.backward(ro, rd, bg, means, scales, rots, opac, shs, D, f["out"], cu(dL).reshape(f["out"].shape),
                         hits=f if use_lists else None, flags=flags, scale_modifier=mod)
        res.update({f"g_{k}": v.cpu().numpy() for k, v in g.items()})
        res["g_opac"] = res["g_opac"].reshape(-1)
    return res


def hit_lists(res):
    """list of per-ray contributing id lists from the (cap, R) layout"""
    cnt = res["hit_cnt"]; g = res["hit_gidx"]
    return [list(g[:min(c, g.shape[0]), r]) for r, c in enumerate(cnt)]


def oracle_lists(f):
    return [list(f["hit_list"][r, :min(c, f["hit_list"].shape[1])]) for r, c in enumerate(f["hit_cnt"])]


def as_dict(sc):
    return dict(means=sc.means, scales=sc.scales, rots=sc.rots, opac=sc.opac, shs=sc.shs)


def test_known_answer_cases(ctx, oracle32):
    g = load_golden("ref_kat.npz")
    for name in g["names"]:
        sc = {k: g[f"{name}/{k}"] for k in ("means", "scales", "rots", "opac", "shs")}
        o, d, D, dL = g[f"{name}/ray_o"], g[f"{name}/ray_d"], int(g[f"{name}/D"]), g[f"{name}/dL"]
        res = run_cuda(ctx, o, d, sc, D, dL)
        assert_close(res["out"], g[f"{name}/out"], REF_ATOL, REF_RTOL, f"{name} forward vs reference")
        assert_close(res["accum_w"], g[f"{name}/accum_w"], 1e-4, 1e-4, f"{name} accum vs reference")
        for k in ("means", "shs", "opac", "scales", "rots"):
            grad_close(res[f"g_{k}"], g[f"{name}/g_{k}"], GRAD_REL, f"{name} d_{k} vs reference")
        f = oracle32.forward(o, d, BG, sc["means"], sc["scales"], sc["rots"], sc["opac"], sc["shs"], D)
        assert hit_lists(res) == oracle_lists(f), f"{name}: contributing hit indices differ from the oracle"
        assert np.array_equal(res["slot_cnt"], f["slot_cnt"]), f"{name}: k-buffer slots differ"
        assert_close(res["out"], f["out"], ORC_ATOL, ORC_RTOL, f"{name} forward vs oracle")


def test_small_scene_vs_reference_golden(ctx):
    g = load_golden("ref_scene_small.npz")
    sc = {k: g[k] for k in ("means", "scales", "rots", "opac", "shs")}
    res = run_cuda(ctx, g["ray_o"], g["ray_d"], sc, int(g["D"]), g["dL"])
    assert_close(res["out"], g["out"], REF_ATOL, REF_RTOL, "forward")
    assert_close(res["accum_w"], g["accum_w"], 1e-4, 1e-4, "accum")
    for k in ("means", "shs", "opac", "scales", "rots"):
        grad_close(res[f"g_{k}"], g[f"g_{k}"], GRAD_REL, f"d_{k}")


def test_config1_forward(ctx, oracle32):
    """BASELINE config #1 (10k Gaussians, 64x64 rays): reference golden + bit-exact hit lists vs the oracle."""
    g = load_golden("ref_cfg1_forward.npz")
    sc = syn.make_street_scene(10000, seed=0)
    o, d = syn.ray_patch(64, 64)
    res = run_cuda(ctx, o, d, as_dict(sc), 3, cap=96)
    assert_close(res["out"], g["out"], 2e-4, REF_RTOL, "forward vs reference")
    assert_close(res["accum_w"], g["accum_w"], 2e-4, 1e-4, "accum vs reference")
    f = oracle32.forward(o, d, BG, sc.means, sc.scales, sc.rots, sc.opac, sc.shs, 3, flags=ORC_BVH, cap=96)
    assert hit_lists(res) == oracle_lists(f)
    assert np.array_equal(res["slot_cnt"], f["slot_cnt"])
    assert_close(res["out"], f["out"], ORC_ATOL, ORC_RTOL, "forward vs oracle")


@pytest.mark.parametrize("P,H,W,seed,D", [(200_000, 16, 512, 5, 3), (50_000, 32, 128, 6, 1)])
def test_mid_scene_vs_oracle(ctx, oracle32, P, H, W, seed, D):
    sc = syn.make_street_scene(P, seed=seed)
    o, d = syn.ray_patch(H, W, frame=2)
    rng = np.random.default_rng(seed)
    dL = np.zeros((H * W, 9), np.float32); dL[:, :4] = rng.standard_normal((H * W, 4))
    res = run_cuda(ctx, o, d, as_dict(sc), D, dL, cap=96)
    args = (o, d, BG, sc.means, sc.scales, sc.rots, sc.opac, sc.shs, D)
    f = oracle32.forward(*args, flags=ORC_BVH, cap=96)
    assert hit_lists(res) == oracle_lists(f), "hit indices must be bit-exact"
    assert np.array_equal(res["slot_cnt"], f["slot_cnt"])
    assert_close(res["out"], f["out"], ORC_ATOL, ORC_RTOL, "forward")
    assert_close(res["accum_w"], f["accum_w"], 1e-4, 1e-4, "accum")
    b = oracle32.backward(*args, f["out"], dL, flags=ORC_BVH)
    for k in ("means", "shs", "opac", "scales", "rots"):
        grad_close(res[f"g_{k}"], b[k], GRAD_REL, f"d_{k}")


def test_backward_list_replay_equals_retrace_and_overflow(ctx):
    sc = syn.make_street_scene(20000, seed=8)
    o, d = syn.ray_patch(32, 64)
    rng = np.random.default_rng(8)
    dL = np.zeros((32 * 64, 9), np.float32); dL[:, :4] = rng.standard_normal((32 * 64, 4))
    a = run_cuda(ctx, o, d, as_dict(sc), 3, dL, cap=128, use_lists=True)
    b = run_cuda(ctx, o, d, as_dict(sc), 3, dL, cap=128, use_lists=False)       # reference-style re-trace
    c = run_cuda(ctx, o, d, as_dict(sc), 3, dL, cap=4, use_lists=True)          # most rays overflow -> re-trace fallback
    assert a["hit_cnt"].max() <= 128 and (c["hit_cnt"] > 4).mean() > 0.5
    for k in ("means", "shs", "opac", "scales", "rots"):
        grad_close(b[f"g_{k}"], a[f"g_{k}"], 1e-4, f"retrace vs list d_{k}")
        grad_close(c[f"g_{k}"], a[f"g_{k}"], 1e-4, f"overflow vs list d_{k}")
    # LRT_FLAG_FIX_BG_GRAD only changes terms proportional to sum_ch dL[ch] bg[ch]
    e = run_cuda(ctx, o, d, as_dict(sc), 3, dL, flags=1)
    assert np.abs(e["g_opac"] - a["g_opac"]).max() > 1e-4


def test_all_forward_kernels_and_options_agree_bitwise(ctx):
    """LRT_OPT_FORWARD_KERNEL 0/1/2/3/4, ray tiles on/off, Morton 30/63: tuning knobs must not change a single bit."""
    from lidar_rt_b200 import native
    sc = syn.make_street_scene(60000, seed=12)
    o, d = syn.ray_patch(32, 96, frame=1)
    ref = None
    try:
        for kernel in (0, 1, 2, 3, 4):
            for tiled in (True, False):
                for morton in (32, 63, 30):
                  for shade in ((0, 1, 2, 3, 4, 5, 6) if kernel >= 3 else (3,)):
                    # shade 3 = split passes, sort + slots fused in one warp-per-ray kernel (default); 6 = the two-kernel form with the
                    # sorted record stream; 4 / 5 = kernels 2 / 3 with the by-length ray ordering switched off
                    ctx.set_option(native.OPT_SPLIT_FUSED, 0 if shade == 6 else 1)
                    ctx.set_option(native.OPT_SORT_RAYS, 0 if shade in (4, 5) else 1)
                    shade = 3 if shade == 6 else (shade - 2 if shade >= 4 else shade)
                    ctx.set_option(native.OPT_FORWARD_KERNEL, kernel)
                    ctx.set_option(native.OPT_MORTON_BITS, morton)
                    ctx.set_option(native.OPT_WAVEFRONT_SHADE, shade)
                    dd = d if tiled else d.reshape(-1, 3)
                    res = run_cuda(ctx, o, dd, as_dict(sc), 3, cap=128)
                    valid = np.arange(res["hit_gidx"].shape[0])[:, None] < res["hit_cnt"][None, :]
                    key = (res["out"], res["hit_cnt"], res["slot_cnt"], np.where(valid, res["hit_gidx"], -1), np.where(valid, res["hit_t"], 0.0),
                           np.where(valid[..., None], res["hit_aux"], 0.0))
                    if ref is None:
                        ref = key
                    else:
                        for a_, b_ in zip(ref, key):
                            assert np.array_equal(a_, b_), f"kernel={kernel} tiled={tiled} morton={morton} shade={shade} differs"
    finally:
        ctx.set_option(native.OPT_FORWARD_KERNEL, 4); ctx.set_option(native.OPT_MORTON_BITS, 32); ctx.set_option(native.OPT_WAVEFRONT_SHADE, 3)
        ctx.set_option(native.OPT_SORT_RAYS, 1); ctx.set_option(native.OPT_SPLIT_FUSED, 1)
    assert_close(res["accum_w"], run_cuda(ctx, o, d, as_dict(sc), 3)["accum_w"], 1e-5, 1e-5, "accum (atomic order)")


def _same_forward(a, b, what):
    valid = np.arange(a["hit_gidx"].shape[0])[:, None] < np.minimum(a["hit_cnt"], a["hit_gidx"].shape[0])[None, :]
    assert np.array_equal(a["out"], b["out"], equal_nan=True), f"{what}: outputs differ"
    assert np.array_equal(a["hit_cnt"], b["hit_cnt"]) and np.array_equal(a["slot_cnt"], b["slot_cnt"]), f"{what}: counts differ"
    assert np.array_equal(np.where(valid, a["hit_gidx"], -1), np.where(valid, b["hit_gidx"], -1)), f"{what}: hit lists differ"


def test_beam_grid_edge_cases_equal_per_ray_traversal(ctx, oracle32):
    """The shared-origin beam grid (kernel 4) against one-thread-per-ray traversal (kernel 0), bit for bit, on the inputs
    that stress its (azimuth, elevation) windows: a tilted sensor, zenith / nadir / duplicate / degenerate directions,
    un-normalised directions, the +-pi azimuth seam, surfels around, above and below the sensor, 1- and 7-ray frames."""
    from lidar_rt_b200 import native
    rng = np.random.default_rng(21)
    sc = syn.make_street_scene(40000, seed=21, scale_mult=2.0)
    o, d = syn.ray_patch(32, 128, frame=2)
    th = 0.6; c, s = np.cos(th), np.sin(th)
    tilt = np.array([[1, 0, 0], [0, c, -s], [0, s, c]], np.float32)
    # big surfels enclosing / next to / right above and below the sensor; one straddling the azimuth seam behind it
    extra = dict(means=np.array([[0.05, 0.02, 0.1], [0.0, 0.0, 1.5], [0.3, -0.2, -1.2], [-6.0, 0.01, 0.0], [2.0, 0.0, 0.0]], np.float32) + o.reshape(1, 3),
                 scales=np.array([[0.4, 0.3], [0.8, 0.8], [0.5, 0.9], [0.7, 0.7], [0.05, 3.0]], np.float32),
                 rots=np.array([[1, 0, 0, 0], [1, 0.02, 0, 0], [1, 0, 0.03, 0], [1, 0, 1, 0], [1, 0.5, 0.5, 0.1]], np.float32),
                 opac=np.full((5, 1), 0.3, np.float32), shs=(0.2 * rng.standard_normal((5, 16, 3))).astype(np.float32))
    scd = {k: np.concatenate([as_dict(sc)[k], extra[k]], 0) for k in extra}
    special = np.array([[0, 0, 1], [0, 0, -1], [0, 0, 1], [1e-9, 0, 1], [-1, 1e-8, 0], [-1, -1e-8, 0], [-1, 0, 0], [0, 0, 0],
                        [3, 0, 0.1], [0, -250.0, 5.0], [1e-4, 1e-4, 1e-5]], np.float32)
    cases = {"upright": d.reshape(-1, 3), "tilted": d.reshape(-1, 3) @ tilt.T,
             "special": np.concatenate([special, d.reshape(-1, 3)[::37] * rng.uniform(0.1, 30, (len(d.reshape(-1, 3)[::37]), 1)).astype(np.float32)], 0),
             "one_ray": d.reshape(-1, 3)[:1], "seven_rays": d.reshape(-1, 3)[100:107]}
    try:
        for name, dd in cases.items():
            dd = np.ascontiguousarray(dd, np.float32)
            ctx.set_option(native.OPT_FORWARD_KERNEL, 0); a = run_cuda(ctx, o, dd, scd, 3, cap=160)
            ctx.set_option(native.OPT_FORWARD_KERNEL, 4); b = run_cuda(ctx, o, dd, scd, 3, cap=160)
            _same_forward(a, b, name)
        # the oracle agrees on the stress scene too (hit indices bit-exact)
        dd = np.ascontiguousarray(cases["tilted"][::5])
        b = run_cuda(ctx, o, dd, scd, 3, cap=160)
        f = oracle32.forward(o, dd, BG, scd["means"], scd["scales"], scd["rots"], scd["opac"], scd["shs"], 3, flags=ORC_BVH, cap=160)
        assert hit_lists(b) == oracle_lists(f) and np.array_equal(b["slot_cnt"], f["slot_cnt"])
        # per-ray origins cannot use the grid: kernel 4 must route them through the hierarchy, same answer
        oo = np.ascontiguousarray(np.broadcast_to(o.reshape(1, 3), (dd.shape[0], 3)) + 0.0)
        c_ = run_cuda(ctx, oo, dd, scd, 3, cap=160)
        _same_forward(b, c_, "per-ray origins")
    finally:
        ctx.set_option(native.OPT_FORWARD_KERNEL, 4)


def test_backward_kernels_agree(ctx, oracle32):
    """thread-per-ray list replay (0), warp-per-ray scans (1), two-pass prefix + thread-per-hit (2, default): same
    gradients up to summation order; the two-pass kernel also with dL/dnormal fed and against the oracle."""
    from lidar_rt_b200 import native
    sc = syn.make_street_scene(60000, seed=13)
    o, d = syn.ray_patch(32, 96, frame=3)
    rng = np.random.default_rng(13)
    dL = np.zeros((32 * 96, 9), np.float32); dL[:, :4] = rng.standard_normal((32 * 96, 4)); dL[:, 5:8] = 0.1 * rng.standard_normal((32 * 96, 3))
    try:
        ctx.set_option(native.OPT_BACKWARD_KERNEL, 0); a = run_cuda(ctx, o, d, as_dict(sc), 3, dL, cap=128)
        ctx.set_option(native.OPT_BACKWARD_KERNEL, 1); b = run_cuda(ctx, o, d, as_dict(sc), 3, dL, cap=128)
        ctx.set_option(native.OPT_BACKWARD_KERNEL, 0); ctx.set_option(native.OPT_SORT_RAYS, 0)
        c = run_cuda(ctx, o, d, as_dict(sc), 3, dL, cap=128)
        ctx.set_option(native.OPT_BACKWARD_KERNEL, 2); ctx.set_option(native.OPT_SORT_RAYS, 1)
        e = run_cuda(ctx, o, d, as_dict(sc), 3, dL, cap=128)
        e1 = run_cuda(ctx, o, d, as_dict(sc), 1, dL, cap=128)                      # SH degree 1: 12-float rows, scalar tail
        e_fix = run_cuda(ctx, o, d, as_dict(sc), 3, dL, cap=128, flags=1)
        ctx.set_option(native.OPT_BACKWARD_KERNEL, 0)
        a1 = run_cuda(ctx, o, d, as_dict(sc), 1, dL, cap=128)
        a_fix = run_cuda(ctx, o, d, as_dict(sc), 3, dL, cap=128, flags=1)
    finally:
        ctx.set_option(native.OPT_BACKWARD_KERNEL, 2); ctx.set_option(native.OPT_SORT_RAYS, 1)
    for k in ("means", "shs", "opac", "scales", "rots"):
        grad_close(b[f"g_{k}"], a[f"g_{k}"], 2e-4, f"d_{k}")
        grad_close(c[f"g_{k}"], a[f"g_{k}"], 2e-4, f"unsorted rays d_{k}")
        grad_close(e[f"g_{k}"], a[f"g_{k}"], 2e-4, f"two-pass d_{k}")
        grad_close(e1[f"g_{k}"], a1[f"g_{k}"], 2e-4, f"two-pass degree 1 d_{k}")
        grad_close(e_fix[f"g_{k}"], a_fix[f"g_{k}"], 2e-4, f"two-pass FIX_BG_GRAD d_{k}")
    ob = oracle32.backward(o, d, BG, sc.means, sc.scales, sc.rots, sc.opac, sc.shs, 3, e["out"], dL, flags=ORC_BVH)
    for k in ("means", "shs", "opac", "scales", "rots"):
        grad_close(e[f"g_{k}"], ob[k], GRAD_REL, f"two-pass vs oracle d_{k}")


def test_refit_matches_rebuild(ctx):
    sc = syn.make_street_scene(30000, seed=9, n_actors=1, per_actor=2000)
    o, d = syn.ray_patch(16, 128)
    base = run_cuda(ctx, o, d, as_dict(sc), 2)
    moved = syn.scene_at_frame(sc, 5)
    fresh = run_cuda(ctx, o, d, as_dict(moved), 2)
    run_cuda(ctx, o, d, as_dict(sc), 2)                       # structure for frame 0 ...
    refit = run_cuda(ctx, o, d, as_dict(moved), 2, refit=True)   # ... refit to frame 5
    assert np.array_equal(refit["out"], fresh["out"]) and hit_lists(refit) == hit_lists(fresh)
    assert not np.array_equal(base["out"], fresh["out"])
    info = ctx.info()
    assert info.refits >= 1 and info.P == 30000


def test_kitti_dynamic_refit_long_bins_vs_oracle(ctx, oracle32):
    """BASELINE config #3 shape at test size: KITTI-360 ray grid (linear inclinations, pixel offset 0), moving actors, structure
    refit to a later frame; large surfels so that many rays carry 65-256+ candidates (the register sorts of k_wf_sort)."""
    H, W = 22, 103
    sc0 = syn.make_street_scene(40_000 + 4 * 2000, seed=21, n_actors=4, per_actor=2000, scale_mult=3.0)
    inc = syn.kitti_inclinations(syn.KITTI_H)[::3]
    o, d = syn.lidar_rays(H, W, inc, syn.sensor_pose(4), pixel_offset=0.0)
    rng = np.random.default_rng(21)
    dL = np.zeros((H * W, 9), np.float32); dL[:, :4] = rng.standard_normal((H * W, 4))
    run_cuda(ctx, o, d, as_dict(sc0), 3)                                     # structure of frame 0 ...
    sc = syn.scene_at_frame(sc0, 4)
    res = run_cuda(ctx, o, d, as_dict(sc), 3, dL, cap=256, refit=True)       # ... refit to frame 4
    assert res["slot_cnt"].max() > 128 and (res["slot_cnt"] > 64).mean() > 0.2, "scene too thin to exercise the long-bin sorts"
    args = (o, d, BG, sc.means, sc.scales, sc.rots, sc.opac, sc.shs, 3)
    f = oracle32.forward(*args, flags=ORC_BVH, cap=256)
    assert hit_lists(res) == oracle_lists(f), "hit indices must be bit-exact"
    assert np.array_equal(res["slot_cnt"], f["slot_cnt"])
    assert_close(res["out"], f["out"], ORC_ATOL, ORC_RTOL, "forward")
    b = oracle32.backward(*args, f["out"], dL, flags=ORC_BVH)
    for k in ("means", "shs", "opac", "scales", "rots"):
        grad_close(res[f"g_{k}"], b[k], GRAD_REL, f"d_{k}")


@pytest.mark.parametrize("n_stack", [700, 3000, 9000])
def test_very_long_candidate_bins_and_bin_overflow(ctx, oracle32, n_stack):
    """Rays that cross hundreds to thousands of surfels (a vehicle's side seen at a grazing angle does this): 700 candidates take
    the in-place global-memory sort, 3000 a 4096-entry bin, 9000 exceed any bin and go to the per-ray fallback. All must equal
    the per-ray traversal kernel bit for bit and the oracle's hit lists."""
    from lidar_rt_b200 import native
    rng = np.random.default_rng(n_stack)
    P = n_stack + 500
    means = np.zeros((P, 3), np.float32)
    means[:n_stack, 0] = 5.0 + 0.002 * np.arange(n_stack); means[:n_stack, 1:] = 0.01 * rng.standard_normal((n_stack, 2))
    means[n_stack:] = rng.uniform(-20, 20, (P - n_stack, 3)) + np.array([30, 0, 0], np.float32)
    scales = np.full((P, 2), 0.4, np.float32)
    rots = np.tile(np.array([np.cos(np.pi / 4), 0, np.sin(np.pi / 4), 0], np.float32), (P, 1))      # normal along +x
    rots[n_stack:] = rng.standard_normal((P - n_stack, 4))
    opac = np.full((P, 1), 0.012, np.float32)
    shs = (0.05 * rng.standard_normal((P, 16, 3))).astype(np.float32); shs[:, 0, :] = 0.5
    sc = dict(means=means, scales=scales, rots=rots, opac=opac, shs=shs)
    yy, zz = np.meshgrid(np.linspace(-0.03, 0.03, 8), np.linspace(-0.03, 0.03, 8), indexing="ij")
    d = np.stack([np.ones_like(yy), yy, zz], -1).astype(np.float32); d /= np.linalg.norm(d, axis=-1, keepdims=True)
    o = np.zeros((1, 3), np.float32)
    res = run_cuda(ctx, o, d, sc, 3, cap=256)
    assert res["slot_cnt"].max() >= 256, "the stack must keep rays alive for hundreds of slots"
    try:
        ctx.set_option(native.OPT_FORWARD_KERNEL, 0)
        ref = run_cuda(ctx, o, d, sc, 3, cap=256)
    finally:
        ctx.set_option(native.OPT_FORWARD_KERNEL, 4)
    _same_forward(res, ref, f"stack of {n_stack}")
    f = oracle32.forward(o, d.reshape(-1, 3), BG, means, scales, rots, opac, shs, 3, flags=ORC_BVH, cap=256)
    assert hit_lists(res) == oracle_lists(f), "hit indices must be bit-exact"
    assert np.array_equal(res["slot_cnt"], f["slot_cnt"])
    assert_close(res["out"], f["out"], ORC_ATOL, ORC_RTOL, "forward")


@pytest.mark.parametrize("fused", [1, 0])
@pytest.mark.parametrize("n_stack,bin_cap", [(700, 512), (3000, 512), (3000, 2048), (9000, 1024), (20000, 512)])
def test_candidates_beyond_the_bin_take_the_overflow_list(ctx, n_stack, bin_cap, fused):
    """A ray with more candidates than its bin holds (LRT_OPT_BIN_CAP forces small bins) must not drop to the per-ray fallback:
    the excess travels through the overflow list into the ray's area of the arena, is sorted by a block (in shared memory up to
    16384 keys, in global memory beyond) and walked by a warp. Outputs, hit lists and slot counts must equal, bit for bit, what the
    same frame gives with bins that hold everything (and, through test_very_long_candidate_bins_and_bin_overflow, the oracle)."""
    import ctypes
    from lidar_rt_b200 import native
    rng = np.random.default_rng(n_stack)
    P = n_stack + 500
    means = np.zeros((P, 3), np.float32)
    means[:n_stack, 0] = 5.0 + (0.002 if n_stack <= 9000 else 0.0009) * np.arange(n_stack); means[:n_stack, 1:] = 0.01 * rng.standard_normal((n_stack, 2))
    means[n_stack:] = rng.uniform(-20, 20, (P - n_stack, 3)) + np.array([30, 0, 0], np.float32)
    scales = np.full((P, 2), 0.4, np.float32)
    rots = np.tile(np.array([np.cos(np.pi / 4), 0, np.sin(np.pi / 4), 0], np.float32), (P, 1))      # normal along +x
    rots[n_stack:] = rng.standard_normal((P - n_stack, 4))
    opac = np.full((P, 1), 0.012 if n_stack <= 9000 else 0.006, np.float32)
    shs = (0.05 * rng.standard_normal((P, 16, 3))).astype(np.float32); shs[:, 0, :] = 0.5
    sc = dict(means=means, scales=scales, rots=rots, opac=opac, shs=shs)
    yy, zz = np.meshgrid(np.linspace(-0.03, 0.03, 8), np.linspace(-0.03, 0.03, 8), indexing="ij")
    d = np.stack([np.ones_like(yy), yy, zz], -1).astype(np.float32); d /= np.linalg.norm(d, axis=-1, keepdims=True)
    o = np.zeros((1, 3), np.float32)
    counters = lambda: (lambda c: (ctx.lib.lrt_debug_counters(ctx._h, c), list(c))[1])((ctypes.c_int * 16)())
    try:
        ctx.set_option(native.OPT_SPLIT_FUSED, fused)
        ctx.set_option(native.OPT_BIN_CAP, 16384)
        ref = run_cuda(ctx, o, d, sc, 3, cap=256)
        c_ref = counters()
        ctx.set_option(native.OPT_BIN_CAP, bin_cap)
        res = run_cuda(ctx, o, d, sc, 3, cap=256)
        c = counters()
    finally:
        ctx.set_option(native.OPT_BIN_CAP, 0); ctx.set_option(native.OPT_SPLIT_FUSED, 1)
    assert res["slot_cnt"].max() >= 256
    assert c[14] > 0 and c[15] > 0, f"the overflow list must have been used (counters {c})"
    if n_stack <= 16384:
        assert c_ref[14] == 0
    assert c[8] == c_ref[8], f"rays handed to the per-ray fallback: {c[8]} with small bins, {c_ref[8]} with large ones"
    if n_stack <= 16384:
        assert c[8] == 0
    _same_forward(res, ref, f"stack of {n_stack}, bins of {bin_cap}")
    if n_stack > 16384:                                   # beyond any bin: the large-bin run used the overflow list too; check against per-ray traversal
        try:
            ctx.set_option(native.OPT_FORWARD_KERNEL, 0)
            ref0 = run_cuda(ctx, o, d, sc, 3, cap=256)
        finally:
            ctx.set_option(native.OPT_FORWARD_KERNEL, 4)
        _same_forward(res, ref0, f"stack of {n_stack} vs per-ray traversal")


@pytest.mark.parametrize("P,H,W,seed", [(200_000, 16, 512, 5), (30_000, 32, 96, 7)])
def test_triangle_depth_mode_vs_triangle_oracle(ctx, oracle32, P, H, W, seed):
    """LRT_OPT_TRIANGLE_DEPTH: hits and depths from the reference's literal proxy (two fp32 triangles per Gaussian, fp64
    Moeller-Trumbore) — the counterpart of the oracle's ORC_TRIANGLES mode, which stands in for what forward.cu:319 reads from OptiX
    (optixGetRayTmax). Hit lists and slot counts bit-exact against that mode; gradients by value."""
    from lidar_rt_b200 import native
    from oracle.oracle import ORC_TRIANGLES
    sc = syn.make_street_scene(P, seed=seed)
    o, d = syn.ray_patch(H, W, frame=2)
    rng = np.random.default_rng(seed)
    dL = np.zeros((H * W, 9), np.float32); dL[:, :4] = rng.standard_normal((H * W, 4))
    try:
        ctx.set_option(native.OPT_TRIANGLE_DEPTH, 1)
        res = run_cuda(ctx, o, d, as_dict(sc), 3, dL, cap=96)
    finally:
        ctx.set_option(native.OPT_TRIANGLE_DEPTH, 0)
    args = (o, d, BG, sc.means, sc.scales, sc.rots, sc.opac, sc.shs, 3)
    f = oracle32.forward(*args, flags=ORC_BVH | ORC_TRIANGLES, cap=96)
    assert hit_lists(res) == oracle_lists(f), "hit indices must be bit-exact against the triangle-mode oracle"
    assert np.array_equal(res["slot_cnt"], f["slot_cnt"])
    assert_close(res["out"], f["out"], ORC_ATOL, ORC_RTOL, "forward")
    b = oracle32.backward(*args, f["out"], dL, flags=ORC_BVH | ORC_TRIANGLES)
    for k in ("means", "shs", "opac", "scales", "rots"):
        grad_close(res[f"g_{k}"], b[k], GRAD_REL, f"d_{k}")
    # and it is a different answer from the analytic quad's on some rays (otherwise the option would be pointless to test)
    res0 = run_cuda(ctx, o, d, as_dict(sc), 3, cap=96)
    assert not np.array_equal(res0["out"][:, 3], res["out"][:, 3])


def test_full_size_properties(ctx):
    """BASELINE config #2 shape (1M Gaussians, 64 x 2650 rays): size-independent invariants."""
    sc = syn.make_street_scene(1_000_000, seed=1)
    o, d = syn.lidar_rays(syn.WAYMO_H, syn.WAYMO_W, syn.waymo_inclinations(), syn.sensor_pose(0))
    R = syn.WAYMO_H * syn.WAYMO_W
    rng = np.random.default_rng(0)
    dL = np.zeros((R, 9), np.float32); dL[:, :4] = rng.standard_normal((R, 4))
    a = run_cuda(ctx, o, d, as_dict(sc), 3, dL, cap=64)
    out = a["out"]
    assert np.isfinite(out).all() and all(np.isfinite(a[f"g_{k}"]).all() for k in ("means", "shs", "opac", "scales", "rots"))
    T, Wsum = out[:, 8], out[:, 4]
    assert_close(Wsum + T, np.ones_like(T), 2e-5, 0, "sum of weights + final transmittance == 1")      # telescoping
    assert (T >= 1e-4 * (1 - 0.99) - 1e-9).all() and (T <= 1).all()
    empty = a["hit_cnt"] == 0
    assert_close(out[empty][:, :3], np.broadcast_to(BG, (int(empty.sum()), 3)), 0, 0, "rays without hits return the background")
    assert_close(float(a["accum_w"].astype(np.float64).sum()), float(Wsum.astype(np.float64).sum()), 0, 1e-5, "checksum of checksums")
    assert (a["slot_cnt"] >= a["hit_cnt"]).all()
    cnt = np.minimum(a["hit_cnt"], 64)
    ts = a["hit_t"]
    for k in range(1, 8):                                         # depth sorted front to back
        m = cnt > k
        assert (ts[k][m] >= ts[k - 1][m]).all()
    # Gaussians never hit get exactly zero gradient; hit ones are the only non-zeros
    touched = np.zeros(sc.P, bool); g = a["hit_gidx"]
    for k in range(64):
        touched[g[k][cnt > k]] = True
    untouched = ~touched & (a["accum_w"] == 0)
    assert (a["g_means"][untouched] == 0).all() and (a["g_opac"][untouched] == 0).all()
    b = run_cuda(ctx, o, d, as_dict(sc), 3, cap=64)               # determinism of the forward
    assert np.array_equal(a["out"], b["out"]) and np.array_equal(a["hit_gidx"][:8], b["hit_gidx"][:8])


def test_autograd_surface_end_to_end():
    """Tracer / raytracing() drop-in: shapes, dict keys, gradients reach leaf parameters."""
    import lib.gaussian_renderer as gr

    class Asset:                                                   # GaussianModel-shaped (gaussian_model.py:112-148)
        def __init__(self, sc):
            dev = "cuda"
            self._xyz = torch.tensor(sc.means, device=dev, requires_grad=True)
            self._scaling = torch.tensor(np.log(sc.scales), device=dev, requires_grad=True)
            self._rotation = torch.tensor(sc.rots, device=dev, requires_grad=True)
            op = np.clip(sc.opac, 1e-4, 1 - 1e-4)
            self._opacity = torch.tensor(np.log(op / (1 - op)), device=dev, requires_grad=True)
            self._features = torch.tensor(sc.shs, device=dev, requires_grad=True)
            self.active_sh_degree = 3
        def get_world_xyz(self, frame): return self._xyz
        @property
        def get_opacity(self): return torch.sigmoid(self._opacity)
        @property
        def get_scaling(self): return torch.exp(self._scaling)
        def get_rotation(self, frame): return torch.zeros((1, 4), device="cuda"), torch.nn.functional.normalize(self._rotation)
        @property
        def get_features(self): return self._features

    sc = syn.make_street_scene(20000, seed=4)
    asset = Asset(sc)
    o, d = syn.ray_patch(16, 64)
    H, W = d.shape[:2]
    centre = torch.tensor(o[0], device="cuda")
    rays_o = centre[None, None].expand(H, W, 3)                    # stride-0 view like LiDARSensor.get_range_rays
    pkg = gr.raytracing(0, [asset], (rays_o, cu(d), centre), torch.tensor([0.0, 0.0, 1.0]), None)
    assert set(pkg) == {"depth", "intensity", "raydrop", "means3D", "accum_gaussian_weight"}
    assert pkg["depth"].shape == (H, W, 1) and pkg["accum_gaussian_weight"].shape == (20000, 1)
    loss = pkg["depth"].mean() + pkg["intensity"].mean() + pkg["raydrop"].mean()
    loss.backward()
    for t in (asset._xyz, asset._scaling, asset._rotation, asset._opacity, asset._features):
        assert t.grad is not None and torch.isfinite(t.grad).all() and t.grad.abs().sum() > 0
    assert pkg["means3D"].grad is None or torch.isfinite(pkg["means3D"].grad).all()


def test_tracer_forward_reads_current_parameters(oracle32):
    """The reference reads the Gaussian parameters on every forward (forward.cu:228-251) and uses the mesh only for its
    BVH: build once, change opacity / scales / means, trace again WITHOUT another build request -> the outputs and the
    gradients must be those of the new parameters (records captured at the last build went stale)."""
    from diff_lidar_tracer import Tracer, TracingSettings
    sc = syn.make_street_scene(30000, seed=12)
    o, d = syn.ray_patch(16, 64)
    H, W = d.shape[:2]
    tracer = Tracer()
    centre = cu(o.reshape(-1)[:3])
    rays_o = centre[None, None].expand(H, W, 3)
    st = TracingSettings(None, None, None, None, cu(BG), 1.0, torch.empty(0, device="cuda"), torch.empty(0, device="cuda"), 3, centre, False, False)
    rng = np.random.default_rng(12)
    dL = np.zeros((H * W, 9), np.float32); dL[:, :4] = rng.standard_normal((H * W, 4))

    def trace(means, scales, rots, opac, request_build):
        ts = [cu(x).requires_grad_(True) for x in (means, scales, rots, opac.reshape(-1, 1), sc.shs)]
        if request_build:
            tracer.build_acceleration_structure(None, None, rebuild=True)
        out, _ = tracer(rays_o, cu(d), None, ts[0], torch.zeros_like(ts[0]), shs=ts[4], opacities=ts[3], scales=ts[1], rotations=ts[2],
                        tracer_settings=st)
        (out.reshape(-1, 9) * cu(dL)).sum().backward()
        return out.detach().reshape(-1, 9).cpu().numpy(), [t.grad.cpu().numpy() for t in ts]

    trace(sc.means, sc.scales, sc.rots, sc.opac, True)
    means2 = (sc.means + rng.normal(0, 0.02, sc.means.shape)).astype(np.float32)
    scales2 = (sc.scales * rng.uniform(0.7, 1.4, sc.scales.shape)).astype(np.float32)
    opac2 = np.clip(sc.opac * rng.uniform(0.5, 1.5, sc.opac.shape), 0.01, 0.999).astype(np.float32)
    out, grads = trace(means2, scales2, sc.rots, opac2, False)
    args = (o, d, BG, means2, scales2, sc.rots, opac2, sc.shs, 3)
    f = oracle32.forward(*args, flags=ORC_BVH, cap=128)
    assert_close(out, f["out"], ORC_ATOL, ORC_RTOL, "forward after a parameter change without a rebuild request")
    b = oracle32.backward(*args, f["out"], dL, flags=ORC_BVH)
    for gt, k in zip(grads, ("means", "scales", "rots", "opac", "shs")):
        grad_close(gt.reshape(b[k].shape), b[k], GRAD_REL, f"d_{k} after a parameter change")
    # assume_static skips the refresh: the caller vouches that nothing changed
    tracer.optix_context.assume_static = True
    out3, _ = trace(means2, scales2, sc.rots, opac2, False)
    assert np.array_equal(out3, out)


# ------------------------------------------------------------------------------------------- edge cases
def _vs_oracle(ctx, oracle32, o, d, sc, D, mod=1.0, with_grad=True, seed=0):
    R = np.asarray(d).reshape(-1, 3).shape[0]
    rng = np.random.default_rng(seed)
    dL = np.zeros((R, 9), np.float32); dL[:, :4] = rng.standard_normal((R, 4))
    res = run_cuda(ctx, o, d, sc, D, dL if with_grad else None, cap=128, mod=mod)
    args = (o, d, BG, sc["means"], sc["scales"], sc["rots"], sc["opac"], sc["shs"], D)
    f = oracle32.forward(*args, flags=ORC_BVH, cap=128, scale_modifier=mod)
    assert hit_lists(res) == oracle_lists(f)
    assert np.array_equal(res["slot_cnt"], f["slot_cnt"])
    assert_close(res["out"], f["out"], ORC_ATOL, ORC_RTOL, "forward")
    if with_grad:
        b = oracle32.backward(*args, f["out"], dL, flags=ORC_BVH, scale_modifier=mod)
        for k in ("means", "shs", "opac", "scales", "rots"):
            grad_close(res[f"g_{k}"], b[k], GRAD_REL, f"d_{k}")
    return res


def test_far_from_world_origin_and_grazing_ground_rays(ctx, oracle32):
    """the sorted-bin scan's fixed millimetre window is not a bound on |t from o - (t' + base)|, which grows like
    ulp(|world coordinate|) / |n.d|. A ground strip seen at grazing angles of 0.3-6 degrees out to 150 m, with scene and sensor
    translated 3 km from the world origin: the windowed paths (beam grid, wavefront) must still give the per-ray traversal's and the
    brute-force oracle's hit lists bit for bit (per-candidate error bounds -> per-ray margins, lrt_trace.cuh)."""
    from lidar_rt_b200 import native
    rng = np.random.default_rng(41)
    P = 9000
    shift = np.array([3000.0, -2000.0, 150.0], np.float32)
    means = np.stack([rng.uniform(4, 150, P), rng.uniform(-2.5, 2.5, P), -2.0 + 0.02 * rng.standard_normal(P)], 1).astype(np.float32)
    nrm = np.tile([0.0, 0.0, 1.0], (P, 1)) + 0.05 * rng.standard_normal((P, 3))
    sc = dict(means=means + shift, scales=np.exp(rng.normal(np.log(0.25), 0.4, (P, 2))).astype(np.float32),
              rots=syn._quat_from_normal(nrm, rng.uniform(0, 2 * np.pi, P)).astype(np.float32),
              opac=np.clip(rng.uniform(0.02, 0.5, (P, 1)), 0.01, 0.999).astype(np.float32), shs=(0.05 * rng.standard_normal((P, 16, 3))).astype(np.float32))
    H, W = 24, 96
    el = -np.radians(np.geomspace(0.3, 6.0, H))[:, None]; az = np.radians(np.linspace(-0.9, 0.9, W))[None, :]
    d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el) * np.ones_like(az)], -1).astype(np.float32)
    o = shift.reshape(1, 3).copy()
    f = oracle32.forward(o, d, BG, sc["means"], sc["scales"], sc["rots"], sc["opac"], sc["shs"], 2, cap=128)      # brute force over all quads
    assert f["slot_cnt"].max() > 32 and f["hit_cnt"].mean() > 4
    try:
        for kernel in (4, 3, 0):
            ctx.set_option(native.OPT_FORWARD_KERNEL, kernel)
            res = run_cuda(ctx, o, d, sc, 2, cap=128)
            assert hit_lists(res) == oracle_lists(f), f"kernel {kernel}: hit indices must be bit-exact"
            assert np.array_equal(res["slot_cnt"], f["slot_cnt"]), f"kernel {kernel}"
            assert_close(res["out"], f["out"], ORC_ATOL, ORC_RTOL, f"kernel {kernel} forward")
    finally:
        ctx.set_option(native.OPT_FORWARD_KERNEL, 4)


def test_per_ray_origins_and_unnormalised_directions(ctx, oracle32):
    """ray_o (R,3) with a different origin per ray (stride 3) and |d| != 1 (t is the ray parameter; SH uses d/|d|)."""
    sc = as_dict(syn.make_street_scene(40000, seed=21))
    _, d = syn.ray_patch(16, 64, frame=1)
    d = d.reshape(-1, 3) * np.linspace(0.5, 2.0, 16 * 64, dtype=np.float32)[:, None]
    rng = np.random.default_rng(21)
    o = (rng.uniform(-1.5, 1.5, (16 * 64, 3)) * np.array([1.0, 1.0, 0.3])).astype(np.float32)
    _vs_oracle(ctx, oracle32, o, d, sc, 2)


def test_scale_modifier(ctx, oracle32):
    sc = as_dict(syn.make_street_scene(30000, seed=22))
    o, d = syn.ray_patch(16, 64)
    a = _vs_oracle(ctx, oracle32, o, d, sc, 3, mod=1.4)
    b = _vs_oracle(ctx, oracle32, o, d, sc, 3, mod=1.0, with_grad=False)
    assert not np.array_equal(a["out"], b["out"])


@pytest.mark.parametrize("P", [1, 7, 8, 9, 65, 513])
def test_tiny_and_ragged_sizes(ctx, oracle32, P):
    """Gaussian counts around the 8-wide node boundaries, single rays, empty ray sets."""
    full = syn.make_street_scene(2000, seed=23, extent=10.0, scale_mult=0.3)
    sc = {k: v[:P].copy() for k, v in as_dict(full).items()}
    sc["means"][:, :2] *= 0.05; sc["means"][:, 2] = np.linspace(2.0, 6.0, P)       # stacked in front of +z
    sc["rots"][:] = np.array([1.0, 0.02, -0.03, 0.0], np.float32)
    o = np.zeros((1, 3), np.float32)
    d = np.array([[0.0, 0.0, 1.0], [0.01, 0.0, 1.0], [0.3, 0.3, 1.0]], np.float32)
    _vs_oracle(ctx, oracle32, o, d, sc, 3)
    _vs_oracle(ctx, oracle32, o, d[:1], sc, 0)
    means, scales, rots, opac, shs = (cu(sc[k]) for k in ("means", "scales", "rots", "opac", "shs"))
    ctx.build(means, scales, rots, opac)
    f = ctx.forward(cu(o), cu(d[:0]), cu(BG), means, scales, rots, opac, shs, 3)          # R = 0
    assert f["out"].shape == (0, 9) and float(f["accum_w"].abs().sum()) == 0.0


def test_invalid_opacities_are_never_hit(ctx, oracle32):
    """opacity < 1/255 -> NaN proxy in the reference (log of < 1): such Gaussians exist but are never hit."""
    sc = as_dict(syn.make_street_scene(20000, seed=24))
    sc["opac"][::3] = 0.002
    o, d = syn.ray_patch(16, 64)
    res = _vs_oracle(ctx, oracle32, o, d, sc, 3)
    assert np.all(res["accum_w"][::3] == 0) and np.all(res["g_opac"][::3] == 0)


def test_error_behaviour(ctx):
    from lidar_rt_b200 import native
    sc = as_dict(syn.make_street_scene(1000, seed=25, extent=10.0))
    means, scales, rots, opac, shs = (cu(sc[k]) for k in ("means", "scales", "rots", "opac", "shs"))
    o, d = syn.ray_patch(4, 8)
    fresh = native.Context()
    with pytest.raises(native.LrtError, match="lrt_build"):                                 # no structure yet
        fresh.forward(cu(o), cu(d), cu(BG), means, scales, rots, opac, shs, 3)
    fresh.build(means, scales, rots, opac)
    with pytest.raises(native.LrtError, match="P differs"):
        fresh.forward(cu(o), cu(d), cu(BG), means[:500], scales[:500], rots[:500], opac[:500], shs[:500], 3)
    with pytest.raises(native.LrtError, match="D <= 3"):
        fresh.forward(cu(o), cu(d), cu(BG), means, scales, rots, opac, shs, 4)
    with pytest.raises(native.LrtError, match="float32"):
        fresh.forward(cu(o), cu(d).double(), cu(BG), means, scales, rots, opac, shs, 3)
    with pytest.raises(native.LrtError, match="num_points, 3"):
        fresh.build(means[None], scales, rots, opac)
    with pytest.raises(native.LrtError, match="refit"):
        native.Context().build(means, scales, rots, opac, refit=True)
    fresh.close()

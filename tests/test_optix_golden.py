"""Parity against the reference ITSELF, run on its real OptiX back end on a B200.

tests/golden/optix_b200.npz and optix_b200_fullsize_sample.npz hold outputs of the unmodified
reference tracer (submodules/diff-lidar-tracer: forward.cu / backward.cu as PTX on OptiX + its `_C`
module), produced on the GPU box by oracle/run_ref_optix.py {golden,bench} from the inputs of
ref_kat.npz / ref_scene_small.npz / config #1 and from the BASELINE workload (P = 2 M, frame 5; `fullsize` mode: forward
on every 16th ray, sparse GRADIENTS for an upstream gradient on every 128th ray; optix_b200_fullsize_1m.npz = config #2, P = 1 M).

What can and cannot match: OptiX's BVH and ray/triangle arithmetic are closed source. Depth `t` of a hit
comes out of that arithmetic, so near-ties order differently and grazing hits move by ~1e-4 relative on a
small fraction of rays. Measured at full size (169 600 rays): median error 1e-6; 99.70 % of rays within
1e-4; 0.034 % beyond 1e-3. The reference's own device code compiled for the host with an fp64 intersector
(oracle/_ref) differs from OptiX on 0.11 % of rays by the same criterion — that is the floor.
"""
import numpy as np
import pytest

from conftest import BG, assert_close, grad_close, load_golden
from lidar_rt_b200 import synthetic as syn
from oracle.oracle import ORC_BVH, ORC_TRIANGLES

KAT_ATOL, KAT_RTOL = 5e-6, 5e-6          # hand-built scenes: no near-ties, no grazing hits
GRAD_REL = 2e-3


def _kat(g, name):
    sc = {k: g[f"{name}/{k}"] for k in ("means", "scales", "rots", "opac", "shs")}
    return sc, g[f"{name}/ray_o"], g[f"{name}/ray_d"], int(g[f"{name}/D"]), g[f"{name}/dL"]


def ray_error(a, b):
    """per-ray error: max abs over (intensity, hit logit, drop logit), depth relative to 1 + |depth|"""
    e = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))
    return np.maximum(e[:, :3].max(1), e[:, 3] / (1.0 + np.abs(b[:, 3])))


# ----------------------------------------------------------------- CPU: the oracle is pinned to real OptiX
@pytest.mark.parametrize("flags", [0, ORC_TRIANGLES], ids=["analytic", "triangles"])
def test_oracle_known_answer_cases_vs_optix(oracle32, flags):
    kat, ox = load_golden("ref_kat.npz"), load_golden("optix_b200.npz")
    for name in kat["names"]:
        sc, o, d, D, dL = _kat(kat, name)
        args = (o, d, BG, sc["means"], sc["scales"], sc["rots"], sc["opac"], sc["shs"], D)
        f = oracle32.forward(*args, flags=flags)
        assert_close(f["out"], ox[f"kat/{name}/out"], KAT_ATOL, KAT_RTOL, f"{name} forward vs OptiX")
        assert_close(f["accum_w"], ox[f"kat/{name}/accum_w"], 1e-5, 1e-5, f"{name} accum vs OptiX")
        b = oracle32.backward(*args, ox[f"kat/{name}/out"], dL, flags=flags)
        for k in ("means", "shs", "opac", "scales", "rots"):
            grad_close(b[k], ox[f"kat/{name}/g_{k}"], GRAD_REL, f"{name} d_{k} vs OptiX")


def test_oracle_small_scene_and_config1_vs_optix(oracle32):
    g, ox = load_golden("ref_scene_small.npz"), load_golden("optix_b200.npz")
    args = (g["ray_o"], g["ray_d"], BG, g["means"], g["scales"], g["rots"], g["opac"], g["shs"], int(g["D"]))
    f = oracle32.forward(*args)
    assert_close(f["out"], ox["small/out"], 5e-5, 1e-5, "small scene forward vs OptiX")
    b = oracle32.backward(*args, ox["small/out"], g["dL"])
    for k in ("means", "shs", "opac", "scales", "rots"):
        grad_close(b[k], ox[f"small/g_{k}"], GRAD_REL, f"small scene d_{k} vs OptiX")
    sc = syn.make_street_scene(10000, seed=0)
    o, d = syn.ray_patch(64, 64)
    f = oracle32.forward(o, d, BG, sc.means, sc.scales, sc.rots, sc.opac, sc.shs, 3, flags=ORC_BVH)
    e = ray_error(f["out"], ox["cfg1/out"])
    assert (e > 1e-4).mean() <= 0.005 and np.median(e) < 1e-5, f"config #1 vs OptiX: {(e > 1e-4).sum()} rays beyond 1e-4"
    assert_close(f["accum_w"], ox["cfg1/accum_w"], 2e-3, 1e-4, "config #1 accum vs OptiX")


def test_host_compiled_reference_goldens_agree_with_optix():
    """oracle/_ref (the fixtures the other tests use) against the real thing."""
    kat, ox = load_golden("ref_kat.npz"), load_golden("optix_b200.npz")
    for name in kat["names"]:
        assert_close(kat[f"{name}/out"], ox[f"kat/{name}/out"], KAT_ATOL, KAT_RTOL, f"{name}")
    assert_close(load_golden("ref_scene_small.npz")["out"], ox["small/out"], 5e-5, 1e-5, "small scene")
    e = ray_error(load_golden("ref_cfg1_forward.npz")["out"], ox["cfg1/out"])
    assert (e > 1e-4).sum() <= 2


def test_oracle_full_size_gradients_vs_optix(oracle32):
    """BASELINE config #2 (P = 1 M, one 64 x 2650 frame): the oracle's outputs and gradients on the golden's gradient rays
    against what the reference produced on OptiX (sparse rows; dL_dout is zero on every other ray, so tracing the subset suffices)."""
    g = load_golden("optix_b200_fullsize_1m.npz")
    P = int(g["P"])
    sc = syn.make_street_scene(P, seed=int(g["seed"]))
    o, d = syn.lidar_rays(64, 2650, syn.waymo_inclinations(), syn.sensor_pose(int(g["frame"])))
    rays = g["grad_rays"]
    dL = np.zeros((len(rays), 9), np.float32)
    dL[:, :4] = np.random.default_rng(int(g["dl_seed"])).standard_normal((len(rays), 4)).astype(np.float32)
    args = (o, d.reshape(-1, 3)[rays], BG, sc.means, sc.scales, sc.rots, sc.opac, sc.shs, 3)
    f = oracle32.forward(*args, flags=ORC_BVH, cap=256)
    e = ray_error(f["out"][:, g["channels"]], g["out_grad_rays"][:, g["channels"]])
    assert np.median(e) < 5e-6 and (e > 1e-4).mean() <= 0.006, f"{(e > 1e-4).sum()} of {len(e)} rays beyond 1e-4"
    b = oracle32.backward(*args, g["out_grad_rays"], dL, flags=ORC_BVH)
    idx = g["g_index"]
    outside = np.ones(P, bool); outside[idx] = False
    for k in ("means", "opac", "scales", "rots"):
        ref = np.asarray(g[f"g_{k}"], np.float64); a = b[k].astype(np.float64)
        err = np.sqrt(np.sum((a[idx] - ref) ** 2) + np.sum(a[outside] ** 2)) / np.linalg.norm(ref)
        assert err <= 2.5e-3, f"d_{k} rel L2 vs OptiX {err:.3e}"
    a = b["shs"][g["g_sh_index"]].astype(np.float64); ref = np.asarray(g["g_shs"], np.float64)
    assert np.linalg.norm(a - ref) / np.linalg.norm(ref) <= 2.5e-3


# ----------------------------------------------------------------- GPU: the CUDA path against real OptiX
@pytest.mark.gpu
def test_cuda_known_answer_and_small_scene_vs_optix():
    from test_gpu_parity import run_cuda
    from lidar_rt_b200 import native
    ctx = native.Context()
    kat, ox = load_golden("ref_kat.npz"), load_golden("optix_b200.npz")
    for name in kat["names"]:
        sc, o, d, D, dL = _kat(kat, name)
        res = run_cuda(ctx, o, d, sc, D, dL)
        assert_close(res["out"], ox[f"kat/{name}/out"], KAT_ATOL, KAT_RTOL, f"{name} forward vs OptiX")
        for k in ("means", "shs", "opac", "scales", "rots"):
            grad_close(res[f"g_{k}"], ox[f"kat/{name}/g_{k}"], GRAD_REL, f"{name} d_{k} vs OptiX")
    g = load_golden("ref_scene_small.npz")
    sc = {k: g[k] for k in ("means", "scales", "rots", "opac", "shs")}
    res = run_cuda(ctx, g["ray_o"], g["ray_d"], sc, int(g["D"]), g["dL"])
    assert_close(res["out"], ox["small/out"], 5e-5, 1e-5, "small scene forward vs OptiX")
    for k in ("means", "shs", "opac", "scales", "rots"):
        grad_close(res[f"g_{k}"], ox[f"small/g_{k}"], GRAD_REL, f"small scene d_{k} vs OptiX")
    ctx.close()


@pytest.mark.gpu
def test_cuda_full_size_vs_optix_sample():
    """BASELINE workload (P = 2 M, 64 x 2650 rays, frame 5): every 16th ray against the reference on OptiX."""
    import torch
    from lidar_rt_b200 import native
    g = load_golden("optix_b200_fullsize_sample.npz")
    sc = syn.make_street_scene(int(g["P"]), seed=int(g["seed"]))
    o, d = syn.lidar_rays(64, 2650, syn.waymo_inclinations(), syn.sensor_pose(int(g["frame"])))
    cu = lambda x: torch.as_tensor(np.ascontiguousarray(x), device="cuda")
    means, scales, rots, opac, shs = map(cu, (sc.means, sc.scales, sc.rots, sc.opac, sc.shs))
    ctx = native.Context()
    ctx.build(means, scales, rots, opac)
    f = ctx.forward(cu(o), cu(d), cu(BG), means, scales, rots, opac, shs, 3)
    out = f["out"].reshape(-1, 9).cpu().numpy()[g["ray_index"]][:, g["channels"]]
    ctx.close()
    e = ray_error(out, g["out"])
    frac4, frac3 = (e > 1e-4).mean(), (e > 1e-3).mean()
    assert np.median(e) < 5e-6, f"median per-ray error {np.median(e):.2e}"
    assert frac4 <= 0.006 and frac3 <= 0.0015, f"{frac4:.4%} of rays beyond 1e-4, {frac3:.4%} beyond 1e-3"
    assert_close(out[:, 4] + out[:, 5], np.ones(len(out)), 2e-5, 0, "accum + final T = 1")


@pytest.mark.gpu
def test_cuda_full_size_vs_optix_sample_triangle_depth():
    """The same frame with LRT_OPT_TRIANGLE_DEPTH = 1: depths from the fp32 proxy triangles instead of the analytic surfel plane.
    DESIGN.md 2 attributed 62 % of the rays beyond 1e-4 against OptiX to that difference; with the option on, what remains is
    OptiX's own closed triangle arithmetic (measured: see profiles/r2_p_triangle_depth.json)."""
    import torch
    from lidar_rt_b200 import native
    g = load_golden("optix_b200_fullsize_sample.npz")
    sc = syn.make_street_scene(int(g["P"]), seed=int(g["seed"]))
    o, d = syn.lidar_rays(64, 2650, syn.waymo_inclinations(), syn.sensor_pose(int(g["frame"])))
    cu = lambda x: torch.as_tensor(np.ascontiguousarray(x), device="cuda")
    means, scales, rots, opac, shs = map(cu, (sc.means, sc.scales, sc.rots, sc.opac, sc.shs))
    ctx = native.Context()
    ctx.build(means, scales, rots, opac)
    fr = {}
    for tri in (0, 1):
        ctx.set_option(native.OPT_TRIANGLE_DEPTH, tri)
        f = ctx.forward(cu(o), cu(d), cu(BG), means, scales, rots, opac, shs, 3, record_hits=False)
        out = f["out"].reshape(-1, 9).cpu().numpy()[g["ray_index"]][:, g["channels"]]
        e = ray_error(out, g["out"])
        fr[tri] = ((e > 1e-4).mean(), (e > 1e-3).mean(), float(np.median(e)))
    ctx.close()
    print("fraction of rays beyond 1e-4 / 1e-3 vs OptiX, median: analytic quad", fr[0], "triangle depth", fr[1])
    # measured on a B200 (round 2): 0.179 % of rays beyond 1e-4 and 0.009 % beyond 1e-3 with the option on, against 0.30 % and
    # 0.034 % for the analytic quad; the reference's own device code with an fp64 intersector (oracle/_ref) sits at 0.11 %
    assert fr[1][0] <= 0.0020 and fr[1][1] <= 0.0003, f"triangle depth: {fr[1][0]:.4%} of rays beyond 1e-4, {fr[1][1]:.4%} beyond 1e-3"
    assert fr[1][0] < 0.7 * fr[0][0] and fr[1][1] < 0.5 * fr[0][1]


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


# full-size gradient parity, by value: BASELINE configs #2 (P = 1 M) and the headline workload (P = 2 M), one 64 x 2650 frame.
# The goldens hold the reference's OptiX gradients for dL_dout = N(0,1) on channels 0-3 of every 128th ray (zero elsewhere),
# stored for the Gaussians they touch (oracle/run_ref_optix.py fullsize). The CUDA path traces the WHOLE frame.
OPTIX_GRAD_REL_L2 = 3.0e-3          # vs the reference on OptiX (closed triangle arithmetic: DESIGN.md 2)
ORACLE_GRAD_REL_L2 = 2e-4           # vs the C oracle on the same rays (same arithmetic; float-atomic order + expf ulps)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["optix_b200_fullsize_sample.npz", "optix_b200_fullsize_1m.npz"], ids=["2M", "1M"])
def test_cuda_full_size_gradients_vs_optix_and_oracle(name, oracle32):
    import torch
    from lidar_rt_b200 import native
    from test_gpu_parity import hit_lists, oracle_lists
    g = load_golden(name)
    P = int(g["P"])
    sc = syn.make_street_scene(P, seed=int(g["seed"]))
    H, W = 64, 2650
    R = H * W
    o, d = syn.lidar_rays(H, W, syn.waymo_inclinations(), syn.sensor_pose(int(g["frame"])))
    grad_rays = g["grad_rays"]
    dL = np.zeros((R, 9), np.float32)
    dL[grad_rays, :4] = np.random.default_rng(int(g["dl_seed"])).standard_normal((len(grad_rays), 4)).astype(np.float32)
    cu = lambda x: torch.as_tensor(np.ascontiguousarray(x), device="cuda")
    means, scales, rots, opac, shs = map(cu, (sc.means, sc.scales, sc.rots, sc.opac, sc.shs))
    ctx = native.Context()
    ctx.build(means, scales, rots, opac)
    f = ctx.forward(cu(o), cu(d), cu(BG), means, scales, rots, opac, shs, 3, want_slots=True)
    gr = ctx.backward(cu(o), cu(d), cu(BG), means, scales, rots, opac, shs, 3, f["out"], cu(dL).reshape(H, W, 9), hits=f)
    out = f["out"].reshape(-1, 9).cpu().numpy()
    cnt_all = f["hit_cnt"].cpu().numpy()
    assert cnt_all.max() <= f["cap"]
    res = dict(hit_cnt=cnt_all[grad_rays], hit_gidx=f["hit_gidx"][:, cu(grad_rays).long()].cpu().numpy())
    slots = f["slot_cnt"].cpu().numpy()[grad_rays]
    G = {k: v.cpu().numpy() for k, v in gr.items()}
    G["opac"] = G["opac"].reshape(-1)
    ctx.close()

    # ---- forward, every 16th ray, against OptiX
    e = ray_error(out[g["ray_index"]][:, g["channels"]], g["out"])
    assert np.median(e) < 5e-6 and (e > 1e-4).mean() <= 0.006 and (e > 1e-3).mean() <= 0.0015

    # ---- gradients against OptiX: everything the reference touched, plus whatever only this path touched
    idx = g["g_index"]
    outside = np.ones(P, bool); outside[idx] = False
    errs = {}
    for k in ("means", "opac", "scales", "rots"):
        ref = np.asarray(g[f"g_{k}"], np.float64)
        a = G[k].astype(np.float64)
        errs[k] = np.sqrt(np.sum((a[idx] - ref) ** 2) + np.sum(a[outside] ** 2)) / np.linalg.norm(ref)
    errs["shs"] = rel_l2(G["shs"][g["g_sh_index"]], g["g_shs"])
    print(f"{name}: gradient rel L2 vs OptiX", {k: f"{v:.3e}" for k, v in errs.items()})
    assert max(errs.values()) <= OPTIX_GRAD_REL_L2, f"{name}: gradient rel L2 vs OptiX {errs}"

    # ---- the same rays through the C oracle: hit lists bit-exact, outputs and gradients by value
    ds = d.reshape(-1, 3)[grad_rays]
    args = (o, ds, BG, sc.means, sc.scales, sc.rots, sc.opac, sc.shs, 3)
    fo = oracle32.forward(*args, flags=ORC_BVH, cap=f["cap"])
    assert hit_lists(res) == oracle_lists(fo), f"{name}: contributing hit indices differ from the oracle"
    assert np.array_equal(slots, fo["slot_cnt"])
    assert_close(out[grad_rays], fo["out"], 2e-5, 2e-5, f"{name}: forward vs oracle")
    bo = oracle32.backward(*args, fo["out"], dL[grad_rays], flags=ORC_BVH)
    for k in ("means", "shs", "opac", "scales", "rots"):
        err = rel_l2(G[k].reshape(bo[k].shape), bo[k])
        assert err <= ORACLE_GRAD_REL_L2, f"{name}: d_{k} rel L2 vs oracle {err:.3e}"
        grad_close(G[k].reshape(bo[k].shape), bo[k], GRAD_REL, f"{name}: d_{k} vs oracle")

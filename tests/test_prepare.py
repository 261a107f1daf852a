"""Fused activation + world transform + concatenation (lrt_prepare) against the reference's own
torch operations (lib/gaussian_renderer/__init__.py:76-134 over the GaussianModel accessors, gaussian_model.py:112-148),
which `lib.gaussian_renderer._assemble` restates op for op. Values within 2e-6 (expf / division ulps), gradients
within 1e-5 of each tensor's max."""
import types

import numpy as np
import pytest
import torch

from conftest import assert_close, grad_close
from lidar_rt_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _assets(n_actors=3, per_actor=500, P_bg=6000, seed=5):
    from lidar_rt_b200.scene import GaussianAsset
    rng = np.random.default_rng(seed)
    sc = syn.make_street_scene(P_bg + n_actors * per_actor, seed=seed, n_actors=n_actors, per_actor=per_actor)
    ids = sc.actor_id
    out = []
    for a in [-1] + list(range(n_actors)):
        m = ids == a
        sub = syn.Scene(sc.means[m], sc.scales[m], sc.rots[m], sc.opac[m], sc.shs[m], ids[m], 3)
        poses = None
        if a >= 0:
            poses = {f: (torch.tensor(rng.uniform(-5, 5, 3), dtype=torch.float32, device="cuda"),
                         torch.tensor(rng.standard_normal((1, 4)) * 1.3, dtype=torch.float32, device="cuda")) for f in (0, 1, 2)}
        out.append(GaussianAsset(sub, device="cuda", actor_poses=poses))
    return out


def _leaf_grads(assets):
    return [[p.grad.detach().cpu().numpy().copy() for p in a.parameters()] for a in assets]


@pytest.mark.parametrize("dynamic,decomp", [(True, False), (True, "object"), (True, "background"), (False, False)])
def test_fused_prepare_matches_torch_path(dynamic, decomp):
    import lib.gaussian_renderer as gr
    from lidar_rt_b200.prepare import fused_prepare
    assets = _assets()
    use = assets[:1] if (decomp == "background" or not dynamic) else (assets[1:] if decomp == "object" else assets)
    frame = 1
    nctx = gr.tracer_2dgs.optix_context.ctx
    rng = np.random.default_rng(0)
    ref = gr._assemble(frame, use, dynamic, decomp)
    ws = [torch.tensor(rng.standard_normal(tuple(t.shape)), dtype=torch.float32, device="cuda") for t in ref]
    sum((t * w).sum() for t, w in zip(ref, ws)).backward()
    g_ref = _leaf_grads(use)
    for a in use:
        for p in a.parameters():
            p.grad = None
    got = fused_prepare(use, frame, dynamic, decomp, nctx)
    assert got is not None, "fused path should apply to GaussianModel-shaped assets"
    for name, a_, b_ in zip(("means3D", "opacity", "scales", "rotations", "shs"), got, ref):
        assert a_.shape == b_.shape, name
        assert_close(a_.detach().cpu().numpy(), b_.detach().cpu().numpy(), 2e-6, 2e-6, name)
    sum((t * w).sum() for t, w in zip(got, ws)).backward()
    g_got = _leaf_grads(use)
    names = ("xyz", "scaling", "rotation", "opacity", "features_dc", "features_rest")
    for k in range(len(use)):
        for nm, x, y in zip(names, g_got[k], g_ref[k]):
            grad_close(x, y, 1e-5, f"asset {k} d_{nm}")


def test_raytracing_with_and_without_fused_prepare():
    """The drop-in render call gives the same picture and the same leaf gradients either way."""
    import lib.gaussian_renderer as gr
    assets = _assets(n_actors=2, per_actor=2000, P_bg=30000, seed=9)
    o, d = syn.ray_patch(16, 96)
    H, W = d.shape[:2]
    centre = torch.tensor(o[0], device="cuda")
    rays = (centre[None, None].expand(H, W, 3), torch.as_tensor(d, device="cuda"), centre)
    bg = torch.tensor([0.0, 0.0, 1.0], device="cuda")
    res = {}
    for fused in (False, True):
        for a in assets:
            for p in a.parameters():
                p.grad = None
        args = types.SimpleNamespace(dynamic=True, pipe=types.SimpleNamespace(fused_prepare=fused, compute_cov3D_python=False, convert_SHs_python=False),
                                     opt=types.SimpleNamespace(use_rayhit=True))
        pkg = gr.raytracing(1, assets, rays, bg, args)
        (pkg["depth"].mean() + pkg["intensity"].mean() + pkg["raydrop"].mean()).backward()
        res[fused] = ({k: pkg[k].detach().cpu().numpy() for k in ("depth", "intensity", "raydrop")}, _leaf_grads(assets))
    for k in ("depth", "intensity", "raydrop"):
        assert_close(res[True][0][k], res[False][0][k], 2e-4, 1e-4, k)
    for k in range(len(assets)):
        for x, y in zip(res[True][1][k], res[False][1][k]):
            grad_close(x, y, 2e-3, f"asset {k}")


@pytest.mark.parametrize("D", [3, 1, 0])
def test_sh_in_place_equals_concatenated_copy(D):
    """raytracing() with the SH coefficients read (and differentiated) in place from features_dc / features_rest
    (lrt_set_sh_parts; pipe.sh_in_place, default) against the same call over the concatenated (P, M, 3) copy: the rendered
    buffers must be bit-identical, the leaf gradients equal up to the order of the float reductions."""
    import lib.gaussian_renderer as gr
    assets = _assets(n_actors=3, per_actor=1500, P_bg=40001, seed=13)         # odd sizes: every alignment of the 180-byte rest rows occurs
    for a in assets:
        a.active_sh_degree = D
    o, d = syn.ray_patch(24, 128)
    H, W = d.shape[:2]
    centre = torch.tensor(o[0], device="cuda")
    rays = (centre[None, None].expand(H, W, 3), torch.as_tensor(d, device="cuda"), centre)
    bg = torch.tensor([0.0, 0.0, 1.0], device="cuda")
    w = torch.randn(H, W, 3, device="cuda")
    res = {}
    for inplace in (False, True):
        for a in assets:
            for p in a.parameters():
                p.grad = None
        args = types.SimpleNamespace(dynamic=True, pipe=types.SimpleNamespace(fused_prepare=True, sh_in_place=inplace, compute_cov3D_python=False, convert_SHs_python=False),
                                     opt=types.SimpleNamespace(use_rayhit=True))
        pkg = gr.raytracing(2, assets, rays, bg, args)
        ((pkg["depth"] * w[..., 0:1]).sum() + (pkg["intensity"] * w[..., 1:2]).sum() + (pkg["raydrop"] * w[..., 2:3]).sum()).backward()
        res[inplace] = ({k: pkg[k].detach().cpu().numpy() for k in ("depth", "intensity", "raydrop")}, _leaf_grads(assets))
    for k in ("depth", "intensity", "raydrop"):
        assert np.array_equal(res[True][0][k], res[False][0][k]), k
    names = ("xyz", "scaling", "rotation", "opacity", "features_dc", "features_rest")
    for k in range(len(assets)):
        for nm, x, y in zip(names, res[True][1][k], res[False][1][k]):
            assert x.shape == y.shape
            grad_close(x, y, 1e-5, f"asset {k} d_{nm}")
    assert np.abs(res[True][1][0][5]).max() > 0 or D == 0


def test_prepare_rejects_bad_input():
    from lidar_rt_b200 import native
    ctx = native.Context()
    a = _assets(n_actors=0, P_bg=100)[0]
    good = dict(xyz=a._xyz.detach(), scaling=a._scaling.detach(), rotation=a._rotation.detach(), opacity=a._opacity.detach(),
                features_dc=a._features_dc.detach(), features_rest=a._features_rest.detach())
    ctx.prepare([good])
    with pytest.raises(native.LrtError):
        ctx.prepare([dict(good, compose_rotation=True)])                      # composed rotation without a pose
    with pytest.raises(native.LrtError):
        ctx.prepare([dict(good, scaling=good["scaling"].double())])
    with pytest.raises(native.LrtError):
        ctx.prepare([])
    ctx.close()

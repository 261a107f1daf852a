"""The step above the tracer pinned to the REFERENCE: tests/golden/ref_prepare.npz holds what the reference's own
statements produce — GaussianModel.get_world_xyz / get_rotation / get_scaling / get_opacity / get_features
(lib/scene/gaussian_model.py:112-148) and the accessor loop, concatenations and rotation composition cut out of raytracing()
(lib/gaussian_renderer/__init__.py:69-134), executed unmodified on the CPU by oracle/make_golden.py — values and leaf gradients,
for a static scene, a dynamic scene (background + 2 actors with poses), decomp="object" and decomp="background"; plus the pinhole
rays of graphics_utils.get_rays for the Camera branch.
  CPU: this repository's accessor path (`_assemble` over scene.GaussianAsset) and `get_rays` against the goldens.
  GPU: lrt_prepare / lrt_prepare_backward (fused_prepare) against the goldens.
"""
import types

import numpy as np
import pytest
import torch

from conftest import assert_close, grad_close, load_golden

NAMES = ("xyz", "scaling", "rotation", "opacity", "features_dc", "features_rest")
CASES = [("static", False, False), ("dynamic", True, False), ("object", True, "object"), ("background", True, "background")]


def _assets(g, device):
    from lidar_rt_b200.scene import GaussianAsset
    out = []
    for k in range(int(g["n_assets"])):
        a = object.__new__(GaussianAsset)
        for nm in NAMES:
            setattr(a, "_" + nm, torch.tensor(g[f"asset{k}/{nm}"], device=device).requires_grad_(True))
        a.active_sh_degree = 3; a.max_sh_degree = 3
        a.actor_poses = None
        if f"asset{k}/pose_T" in g.files:
            a.actor_poses = {1: (torch.tensor(g[f"asset{k}/pose_T"], device=device), torch.tensor(g[f"asset{k}/pose_quat"], device=device))}
        out.append(a)
    return out


def _check(g, tag, assets, use, res):
    for n, t in zip(("means3D", "opacity", "scales", "rotations", "shs"), res):
        assert tuple(t.shape) == g[f"{tag}/{n}"].shape, n
        assert_close(t.detach().cpu().numpy(), g[f"{tag}/{n}"], 2e-6, 2e-6, f"{tag} {n} vs reference")
    for a in use:
        for p in a.parameters():
            p.grad = None
    ws = [torch.tensor(g[f"{tag}/w_{n}"], device=res[0].device) for n in ("means3D", "opacity", "scales", "rotations", "shs")]
    sum((t * w).sum() for t, w in zip(res, ws)).backward()
    for a in use:
        k = assets.index(a)
        for nm, p in zip(NAMES, a.parameters()):
            grad_close(p.grad.cpu().numpy(), g[f"{tag}/g_asset{k}/{nm}"], 1e-5, f"{tag} asset {k} d_{nm} vs reference")


@pytest.mark.parametrize("tag,dynamic,decomp", CASES)
def test_accessor_path_matches_reference_statements(tag, dynamic, decomp):
    import lib.gaussian_renderer as gr
    g = load_golden("ref_prepare.npz")
    assets = _assets(g, "cpu")
    use = [assets[i] for i in g[f"{tag}/assets"]]
    _check(g, tag, assets, use, gr._assemble(1, use, dynamic, decomp))


def test_pinhole_rays_match_reference_get_rays():
    import lib.gaussian_renderer as gr
    g = load_golden("ref_prepare.npz")
    ro, rd = gr.get_rays(g["cam/K"], torch.tensor(g["cam/c2w"]))
    assert_close(rd.numpy(), g["cam/rays_d"], 1e-6, 1e-6, "rays_d")
    assert np.array_equal(ro.contiguous().numpy(), g["cam/rays_o"]) and ro.stride()[:2] == (0, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,dynamic,decomp", CASES)
def test_fused_prepare_matches_reference_statements(tag, dynamic, decomp):
    import lib.gaussian_renderer as gr
    from lidar_rt_b200.prepare import fused_prepare
    g = load_golden("ref_prepare.npz")
    assets = _assets(g, "cuda")
    use = [assets[i] for i in g[f"{tag}/assets"]]
    got = fused_prepare(use, 1, dynamic, decomp, gr.tracer_2dgs.optix_context.ctx)
    assert got is not None, "the fused path applies to GaussianModel-shaped assets"
    _check(g, tag, assets, use, got)


@pytest.mark.gpu
def test_raytracing_camera_sensor_branch(oracle32):
    """raytracing() with a pinhole Camera (reference :31-41): rays from get_rays (un-normalised directions, one shared origin),
    checked against the oracle on the same rays."""
    import lib.gaussian_renderer as gr
    from conftest import BG
    from lidar_rt_b200 import synthetic as syn
    from lidar_rt_b200.scene import GaussianAsset
    from oracle.oracle import ORC_BVH
    sc = syn.make_street_scene(40000, seed=31)
    asset = GaussianAsset(sc, device="cuda")
    c2w = np.eye(4, dtype=np.float32); c2w[:3, :3] = np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]], np.float32); c2w[:3, 3] = (2.0, 0.5, -0.3)
    w2c = np.linalg.inv(c2w)
    cam = types.SimpleNamespace(image_width=48, image_height=32, FoVx=1.3, world_view_transform=torch.tensor(w2c.T, device="cuda"),
                                camera_center=torch.tensor(c2w[:3, 3], device="cuda"))
    pkg = gr.raytracing(0, [asset], cam, torch.tensor(BG), None)
    assert pkg["depth"].shape == (32, 48, 1)
    ro, rd = gr.get_rays([[0.5 * 48 / np.tan(0.65), 0, 24.0], [0, 0.5 * 48 / np.tan(0.65), 16.0], [0, 0, 1]], torch.tensor(c2w[:3, :4], device="cuda"))
    f = oracle32.forward(c2w[:3, 3].reshape(1, 3), rd.cpu().numpy(), BG, sc.means, sc.scales, sc.rots, sc.opac, sc.shs, 3, flags=ORC_BVH, cap=128)
    assert_close(pkg["depth"].reshape(-1).detach().cpu().numpy(), f["out"][:, 3], 1e-4, 1e-4, "camera depth vs oracle")
    assert_close(pkg["intensity"].reshape(-1).detach().cpu().numpy(), f["out"][:, 0], 1e-4, 1e-4, "camera intensity vs oracle")
    assert float(pkg["depth"].max()) > 1.0

"""Range-image ray generation and back-projection kernels (lrt_range_rays / lrt_range_points) against
golden vectors produced by the reference's own LiDARSensor.get_range_rays / range2point (tests/golden/ref_python.npz,
ref_range2point.npz; generators: oracle/make_golden.py, oracle/make_golden_rays.py). fp32 sin/cos of angles up to pi:
2e-6 on unit directions, 2e-4 abs on points up to 80 m away (5e-6 relative)."""
import numpy as np
import pytest
import torch

from conftest import assert_close, load_golden
from lidar_rt_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from lidar_rt_b200 import native
    c = native.Context()
    yield c
    c.close()


@pytest.mark.parametrize("tag,off", [("waymo", 0.5), ("kitti", 0.0)])
def test_range_rays_match_reference(ctx, tag, off):
    g = load_golden("ref_python.npz")
    want_o, want_d = g[f"rays_{tag}_o"], g[f"rays_{tag}_d"]
    H, W = want_d.shape[:2]
    ro, rd = ctx.range_rays(H, W, g[f"rays_{tag}_inc"], g[f"rays_{tag}_pose"], pixel_offset=off)
    assert ro.shape == (H, W, 3) and ro.stride()[:2] == (0, 0)              # the reference's expanded view
    assert_close(rd.cpu().numpy(), want_d, 2e-6, 0, f"{tag} directions")
    assert_close(ro.cpu().numpy(), want_o.reshape(H, W, 3), 0, 0, f"{tag} origins")


@pytest.mark.parametrize("tag", ["waymo", "kitti"])
def test_range_points_match_reference(ctx, tag):
    g = load_golden("ref_range2point.npz")
    off, aoff = (float(v) for v in g[f"{tag}_offsets"])
    p = ctx.range_points(torch.as_tensor(g[f"{tag}_range"], device="cuda"), g[f"{tag}_inc"], g[f"{tag}_pose"], pixel_offset=off, angle_offset=aoff)
    assert_close(p.cpu().numpy(), g[f"{tag}_points"], 2e-4, 5e-6, f"{tag} points")


def test_generated_rays_drive_the_tracer(ctx):
    """Full-size Waymo grid from the kernel == the host generator used everywhere else; traced through the stride-0 origin."""
    from lidar_rt_b200 import native
    pose = syn.sensor_pose(3)
    ro, rd = ctx.range_rays(syn.WAYMO_H, syn.WAYMO_W, syn.waymo_inclinations(), pose)
    o, d = syn.lidar_rays(syn.WAYMO_H, syn.WAYMO_W, syn.waymo_inclinations(), pose)
    assert_close(rd.cpu().numpy(), d, 2e-6, 0, "directions vs host generator")
    sc = syn.make_street_scene(30000, seed=2)
    cu = lambda x: torch.as_tensor(np.ascontiguousarray(x), device="cuda")
    means, scales, rots, opac, shs = map(cu, (sc.means, sc.scales, sc.rots, sc.opac, sc.shs))
    ctx.build(means, scales, rots, opac)
    f = ctx.forward(ro, rd, cu(np.array([0, 0, 1], np.float32)), means, scales, rots, opac, shs, 3)
    out = f["out"].reshape(-1, 9)
    assert torch.isfinite(out).all() and float(out[:, 4].max()) > 0.5
    # back-projecting the rendered expected depth of fully opaque rays lands on the ray
    pts = ctx.range_points(out[:, 3].reshape(syn.WAYMO_H, syn.WAYMO_W).contiguous(), syn.waymo_inclinations(), pose)
    along = ((pts - cu(o).reshape(1, 1, 3)) * rd).sum(-1).reshape(-1)
    assert_close(along.cpu().numpy(), out[:, 3].cpu().numpy(), 2e-4, 1e-5, "range2point o rays")
    with pytest.raises(native.LrtError):
        ctx.range_rays(8, 16, np.zeros(5, np.float32), pose)

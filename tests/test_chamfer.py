"""Chamfer distance: lrt_chamfer_forward / lrt_chamfer_backward against
 * tests/golden/chamfer_ref_b200.npz — outputs of the UNMODIFIED reference extension (lib/utils/chamfer3D) run on a B200
   by oracle/run_ref_chamfer.py golden;
 * the C restatement oracle/chamfer_oracle.c (itself pinned to the same golden file on the CPU);
 * size-independent properties at LiDAR-frame size.
Bar: distances and indices BIT-EXACT (fp32 arithmetic restated operation for operation; lowest index on ties);
gradients within 1e-6 of the largest entry (float atomics commute but do not associate)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, grad_close
from oracle.run_ref_chamfer import case_grads, chamfer_cases, lidar_clouds

GOLD = os.path.join(GOLDEN, "chamfer_ref_b200.npz")
GRAD_REL = 1e-6


@pytest.fixture(scope="module")
def chm():
    from oracle.oracle import ChamferOracle
    return ChamferOracle()


@pytest.fixture(scope="module")
def gold():
    assert os.path.exists(GOLD), "tests/golden/chamfer_ref_b200.npz missing (oracle/run_ref_chamfer.py golden, on a GPU box)"
    return np.load(GOLD)


# ------------------------------------------------------------------------------------------ CPU: the oracle
def test_golden_inputs_are_the_seeded_cases(gold):
    for name, (a, c) in chamfer_cases().items():
        assert np.array_equal(gold[f"{name}/a"], a) and np.array_equal(gold[f"{name}/c"], c), name


def test_oracle_vs_reference_golden(chm, gold):
    for name, (a, c) in chamfer_cases().items():
        d1, d2, i1, i2 = chm.forward(a, c)
        assert np.array_equal(d1, gold[f"{name}/d1"]) and np.array_equal(d2, gold[f"{name}/d2"]), f"{name}: distances not bit-exact"
        assert np.array_equal(i1, gold[f"{name}/i1"]) and np.array_equal(i2, gold[f"{name}/i2"]), f"{name}: indices differ"
        ga, gc = chm.backward(a, c, gold[f"{name}/g1"], gold[f"{name}/g2"], i1, i2)
        grad_close(ga, gold[f"{name}/ga"], GRAD_REL, f"{name} grad_xyz1")
        grad_close(gc, gold[f"{name}/gc"], GRAD_REL, f"{name} grad_xyz2")


def test_oracle_vs_fp64_brute_force_and_tie_rule(chm):
    rng = np.random.default_rng(3)
    a = rng.normal(size=(1, 1200, 3)).astype(np.float32); c = rng.normal(size=(1, 900, 3)).astype(np.float32)
    d1, d2, i1, i2 = chm.forward(a, c)
    D = ((c[0][None, :, :].astype(np.float64) - a[0][:, None, :]) ** 2).sum(-1)
    assert np.allclose(d1[0], D.min(1), rtol=1e-6, atol=1e-12) and np.allclose(d2[0], D.min(0), rtol=1e-6, atol=1e-12)
    assert (i1[0] == D.argmin(1)).mean() > 0.999 and (i2[0] == D.argmin(0)).mean() > 0.999
    # exact ties: every lattice point occurs many times; the first occurrence must be reported
    a = rng.integers(0, 4, (1, 500, 3)).astype(np.float32); c = rng.integers(0, 4, (1, 700, 3)).astype(np.float32)
    d1, _, i1, _ = chm.forward(a, c)
    D = ((c[0][None] - a[0][:, None]) ** 2).sum(-1)
    assert np.array_equal(i1[0], D.argmin(1)) and np.array_equal(d1[0], D.min(1))      # numpy argmin = first occurrence


def test_oracle_vs_independent_kd_tree(chm):
    """An implementation that shares nothing with the oracle (scipy's cKDTree, fp64): same nearest distances on LiDAR-shaped clouds;
    same neighbour wherever the nearest distance is not tied within fp32 resolution."""
    from scipy.spatial import cKDTree
    pr, gt = lidar_clouds(20000, seed=4)
    d1, d2, i1, i2 = chm.forward(pr[None], gt[None])
    for (q, t, d, i) in ((pr, gt, d1[0], i1[0]), (gt, pr, d2[0], i2[0])):
        dist, idx = cKDTree(t.astype(np.float64)).query(q.astype(np.float64), k=2)
        assert np.allclose(d, dist[:, 0] ** 2, rtol=2e-5, atol=1e-10)
        clear = dist[:, 1] ** 2 - dist[:, 0] ** 2 > 1e-5 * (1.0 + dist[:, 0] ** 2)      # runner-up clearly farther
        assert clear.mean() > 0.9 and np.array_equal(i[clear], idx[clear, 0])


def test_oracle_backward_is_the_gradient(chm):
    rng = np.random.default_rng(4)
    a = rng.normal(size=(2, 300, 3)).astype(np.float32); c = rng.normal(size=(2, 200, 3)).astype(np.float32)
    d1, d2, i1, i2 = chm.forward(a, c)
    g1 = rng.normal(size=d1.shape).astype(np.float32); g2 = rng.normal(size=d2.shape).astype(np.float32)
    ga, gc = chm.backward(a, c, g1, g2, i1, i2)
    ta = torch.tensor(a, dtype=torch.float64, requires_grad=True); tc = torch.tensor(c, dtype=torch.float64, requires_grad=True)
    e1 = ((ta - torch.gather(tc, 1, torch.tensor(i1, dtype=torch.int64)[..., None].expand(-1, -1, 3))) ** 2).sum(-1)
    e2 = ((tc - torch.gather(ta, 1, torch.tensor(i2, dtype=torch.int64)[..., None].expand(-1, -1, 3))) ** 2).sum(-1)
    ((e1 * torch.tensor(g1)).sum() + (e2 * torch.tensor(g2)).sum()).backward()
    grad_close(ga, ta.grad.numpy(), 1e-5, "grad_xyz1"); grad_close(gc, tc.grad.numpy(), 1e-5, "grad_xyz2")


def test_oracle_empty_clouds(chm):
    a = np.zeros((1, 4, 3), np.float32) + 1; c = np.zeros((1, 0, 3), np.float32)
    d1, d2, i1, i2 = chm.forward(a, c)
    assert d1.shape == (1, 4) and not d1.any() and not i1.any() and d2.shape == (1, 0)


def test_python_surface_matches_reference():
    import inspect
    from lib.utils.chamfer3D import dist_chamfer_3D as m
    assert issubclass(m.chamfer_3DDist, torch.nn.Module) and issubclass(m.chamfer_3DFunction, torch.autograd.Function)
    assert list(inspect.signature(m.chamfer_3DDist.forward).parameters) == ["self", "input1", "input2"]
    assert list(inspect.signature(m.chamfer_3DFunction.forward).parameters) == ["ctx", "xyz1", "xyz2"]
    assert list(inspect.signature(m.chamfer_3DFunction.backward).parameters) == ["ctx", "graddist1", "graddist2", "gradidx1", "gradidx2"]


# ------------------------------------------------------------------------------------------ GPU: the CUDA path
@pytest.fixture(scope="module")
def ctx():
    from lidar_rt_b200 import native
    c = native.Context("cuda:0")
    yield c
    c.close()


def _cu(x, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(x), device="cuda:0") if dtype is None else torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device="cuda:0")


def _run(ctx, a, c, g1=None, g2=None):
    ta, tc = _cu(a), _cu(c)
    d1, d2, i1, i2 = ctx.chamfer_forward(ta, tc)
    out = [t.cpu().numpy() for t in (d1, d2, i1, i2)]
    if g1 is not None:
        ga, gc = ctx.chamfer_backward(ta, tc, _cu(g1), _cu(g2), i1, i2)
        out += [ga.cpu().numpy(), gc.cpu().numpy()]
    return out


@pytest.mark.gpu
def test_cuda_vs_reference_golden(ctx, gold):
    for name, (a, c) in chamfer_cases().items():
        d1, d2, i1, i2, ga, gc = _run(ctx, a, c, gold[f"{name}/g1"], gold[f"{name}/g2"])
        assert np.array_equal(d1, gold[f"{name}/d1"]) and np.array_equal(d2, gold[f"{name}/d2"]), f"{name}: distances not bit-exact"
        assert np.array_equal(i1, gold[f"{name}/i1"]) and np.array_equal(i2, gold[f"{name}/i2"]), f"{name}: indices differ"
        grad_close(ga, gold[f"{name}/ga"], GRAD_REL, f"{name} grad_xyz1")
        grad_close(gc, gold[f"{name}/gc"], GRAD_REL, f"{name} grad_xyz2")


@pytest.mark.gpu
@pytest.mark.parametrize("n,m,seed,kind", [(20000, 15000, 0, "normal"), (4097, 30001, 1, "uniform"), (30000, 30000, 2, "lattice"),
                                           (25000, 25000, 3, "lidar"), (1000, 64, 4, "line")])
def test_cuda_vs_oracle_seeded(ctx, chm, n, m, seed, kind):
    rng = np.random.default_rng(seed)
    if kind == "normal":
        a = rng.normal(size=(1, n, 3)).astype(np.float32) * 10; c = rng.normal(size=(1, m, 3)).astype(np.float32) * 10
    elif kind == "uniform":
        a = rng.uniform(-50, 50, (1, n, 3)).astype(np.float32); c = rng.uniform(-50, 50, (1, m, 3)).astype(np.float32)
    elif kind == "lattice":      # exact ties everywhere
        a = rng.integers(0, 12, (1, n, 3)).astype(np.float32); c = rng.integers(0, 12, (1, m, 3)).astype(np.float32)
    elif kind == "line":         # degenerate extent: all points on one axis, duplicates
        a = np.zeros((1, n, 3), np.float32); a[0, :, 0] = rng.integers(0, 50, n)
        c = np.zeros((1, m, 3), np.float32); c[0, :, 0] = rng.integers(0, 50, m)
    else:
        pr, gt = lidar_clouds(n, seed=seed)
        a, c = pr[None], gt[None]
    g1, g2 = case_grads(kind, 1, a.shape[1], c.shape[1])
    d1, d2, i1, i2, ga, gc = _run(ctx, a, c, g1, g2)
    o1, o2, j1, j2 = chm.forward(a, c)
    assert np.array_equal(d1, o1) and np.array_equal(d2, o2), "distances not bit-exact vs the oracle"
    assert np.array_equal(i1, j1) and np.array_equal(i2, j2), "indices differ from the oracle"
    oa, oc = chm.backward(a, c, g1, g2, j1, j2)
    grad_close(ga, oa, GRAD_REL, "grad_xyz1"); grad_close(gc, oc, GRAD_REL, "grad_xyz2")


@pytest.mark.gpu
def test_cuda_full_frame_vs_oracle_and_properties(ctx, chm):
    """One Waymo frame's worth of points (the train.py:197-207 call): bit-exact against the O(n·m) oracle, plus
    properties that need no oracle."""
    pr, gt = lidar_clouds(169600, seed=11)
    a, c = pr[None], gt[None]
    ta, tc = _cu(a), _cu(c)
    d1, d2, i1, i2 = ctx.chamfer_forward(ta, tc)
    # properties: indices in range; the distance is the distance to the reported neighbour; no closer point among a sample
    assert int(i1.min()) >= 0 and int(i1.max()) < c.shape[1] and int(i2.min()) >= 0 and int(i2.max()) < a.shape[1]
    x = tc[0][i1[0].long()] - ta[0]
    rec = torch.addcmul(torch.addcmul(x[:, 1] * x[:, 1], x[:, 0], x[:, 0]), x[:, 2], x[:, 2])
    assert torch.allclose(rec, d1[0], rtol=1e-6, atol=0)
    sub = torch.arange(0, a.shape[1], 97, device="cuda:0")
    brute = torch.cdist(ta[0][sub].double(), tc[0].double()).min(1).values ** 2
    assert torch.allclose(brute.float(), d1[0][sub], rtol=1e-5, atol=1e-9)
    # chamfer of a cloud with itself is zero and, without duplicates, every point is its own neighbour
    s1, s2, k1, k2 = ctx.chamfer_forward(ta, ta.clone())
    assert float(s1.abs().max()) == 0.0 and float(s2.abs().max()) == 0.0
    assert torch.equal(k1, k2)
    o1, o2, j1, j2 = chm.forward(a, c)
    assert np.array_equal(d1.cpu().numpy(), o1) and np.array_equal(d2.cpu().numpy(), o2)
    assert np.array_equal(i1.cpu().numpy(), j1) and np.array_equal(i2.cpu().numpy(), j2)


@pytest.mark.gpu
def test_cuda_batches_and_empty_clouds(ctx, chm):
    rng = np.random.default_rng(9)
    a = rng.normal(size=(3, 513, 3)).astype(np.float32); c = rng.normal(size=(3, 77, 3)).astype(np.float32)
    d1, d2, i1, i2 = _run(ctx, a, c)
    o1, o2, j1, j2 = chm.forward(a, c)
    assert np.array_equal(d1, o1) and np.array_equal(d2, o2) and np.array_equal(i1, j1) and np.array_equal(i2, j2)
    e = np.zeros((1, 0, 3), np.float32)
    d1, d2, i1, i2 = _run(ctx, a[:1], e)
    assert d1.shape == (1, 513) and not d1.any() and not i1.any() and d2.shape == (1, 0)
    ga, gc = ctx.chamfer_backward(_cu(a[:1]), _cu(e), _cu(np.ones((1, 513), np.float32)), _cu(np.zeros((1, 0), np.float32)),
                                  _cu(i1), _cu(i2))
    assert not ga.cpu().numpy().any() and gc.shape == (1, 0, 3)


@pytest.mark.gpu
def test_drop_in_module_autograd_like_train_py(chm):
    """train.py:197-207: chamLoss(pred[None], gt[None]) -> (dist1 + dist2).mean() * 0.5 -> backward."""
    from lib.utils.chamfer3D.dist_chamfer_3D import chamfer_3DDist
    pr, gt = lidar_clouds(16384, seed=2)
    p = _cu(pr).requires_grad_(True)
    g = _cu(gt)
    chamLoss = chamfer_3DDist()
    dist1, dist2, idx1, idx2 = chamLoss(p[None, ...], g[None, ...])
    assert idx1.dtype == torch.int32 and dist1.shape == (1, pr.shape[0]) and dist2.shape == (1, gt.shape[0])
    loss = (dist1 + dist2).mean() * 0.5
    loss.backward()
    n = pr.shape[0]
    o1, o2, j1, j2 = chm.forward(pr[None], gt[None])
    w = np.full((1, n), 0.5 / n, np.float32)
    oa, _ = chm.backward(pr[None], gt[None], w, w, j1, j2)
    assert abs(float(loss) - float((o1 + o2).mean() * 0.5)) <= 1e-6 * float(loss)
    grad_close(p.grad.cpu().numpy(), oa[0], GRAD_REL, "d loss / d pred_pts")


@pytest.mark.gpu
def test_error_behaviour(ctx):
    from lidar_rt_b200 import native
    a = torch.zeros((1, 8, 3), device="cuda:0")
    with pytest.raises(native.LrtError):
        ctx.chamfer_forward(a.cpu(), a)
    with pytest.raises(native.LrtError):
        ctx.chamfer_forward(a.double(), a)
    with pytest.raises(native.LrtError):
        ctx.chamfer_forward(a[0], a[0])
    with pytest.raises(native.LrtError):
        ctx.chamfer_forward(a, torch.zeros((2, 8, 3), device="cuda:0"))
    with pytest.raises(native.LrtError):
        ctx.chamfer_forward(torch.zeros((1, 8, 2), device="cuda:0"), a)

"""Densify / prune through the native row-compaction kernels (lrt_compact_rows, lrt_densify_rows) against the
REFERENCE's own methods: tests/golden/ref_densify.npz holds what GaussianModel.prune_points and GaussianModel.densify_and_prune
(lib/scene/gaussian_model.py:235-407, with _prune_optimizer / cat_tensors_to_optimizer / densification_postfix /
densify_and_clone / densify_and_split) leave behind — parameters, both Adam moments, step counts, statistics — when run
unmodified on a seeded 400-Gaussian model with a real torch.optim.Adam (oracle/make_golden.py densify). The split's normal
samples were recorded there and are fed to the kernel here.
Bar: surviving rows index-exact and moved BIT-exactly (parameters and moments); the two computed fields of the split children
(position = R(q) sample + xyz, scale = log(exp(s) / 1.6)) within 2e-6 (expf / logf ulps, product order of the 3x3 by 3x1)."""
import types

import numpy as np
import pytest
import torch

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu

NAMES = [("xyz", "_xyz"), ("f_dc", "_features_dc"), ("f_rest", "_features_rest"), ("opacity", "_opacity"), ("scaling", "_scaling"), ("rotation", "_rotation")]


class Model:
    """GaussianModel-shaped (gaussian_model.py:25-60, :186-201)."""

    def __init__(self, g, opt_cls):
        dev = "cuda"
        for n, a in NAMES:
            setattr(self, a, torch.nn.Parameter(torch.tensor(g[f"in/{n}"], device=dev)))
        self.optimizer = opt_cls([{"params": [getattr(self, a)], "lr": 1e-3 * (k + 1), "name": n} for k, (n, a) in enumerate(NAMES)], lr=0.0, eps=1e-15)
        for n, a in NAMES:
            p = getattr(self, a)
            self.optimizer.state[p] = {"step": torch.tensor(float(g[f"in/{n}/step"])), "exp_avg": torch.tensor(g[f"in/{n}/exp_avg"], device=dev),
                                       "exp_avg_sq": torch.tensor(g[f"in/{n}/exp_avg_sq"], device=dev)}
        self.xyz_gradient_accum = torch.tensor(g["in/xyz_gradient_accum"], device=dev); self.denom = torch.tensor(g["in/denom"], device=dev)
        self.max_radii2D = torch.tensor(g["in/max_radii2D"], device=dev)
        o = g["opt"]
        self.densify_scale_threshold, self.extent, self.dimension, self.bounding_box = float(o[3]), float(o[4]), 2, None

    get_scaling = property(lambda self: torch.exp(self._scaling))
    get_opacity = property(lambda self: torch.sigmoid(self._opacity))


def _compare(model, g, tag, exact_rows=None):
    for n, a in NAMES:
        p = getattr(model, a)
        st = model.optimizer.state[p]
        assert model.optimizer.param_groups[[x for x, _ in NAMES].index(n)]["params"][0] is p
        want = g[f"{tag}/{n}"]
        got = p.detach().cpu().numpy()
        assert got.shape == want.shape, f"{tag} {n}: {got.shape} vs {want.shape}"
        if n in ("xyz", "scaling") and exact_rows is not None:
            assert np.array_equal(got[:exact_rows], want[:exact_rows]), f"{tag} {n}: moved rows must be bit-exact"
            assert_close(got[exact_rows:], want[exact_rows:], 2e-6, 2e-6, f"{tag} {n}: split children")
        else:
            assert np.array_equal(got, want), f"{tag} {n}: rows must be moved bit-exactly"
        for k in ("exp_avg", "exp_avg_sq"):
            assert np.array_equal(st[k].cpu().numpy(), g[f"{tag}/{n}/{k}"]), f"{tag} {n} {k}"
        assert float(st["step"]) == float(g[f"{tag}/{n}/step"])
    for nm in ("xyz_gradient_accum", "denom", "max_radii2D"):
        assert np.array_equal(getattr(model, nm).cpu().numpy(), g[f"{tag}/{nm}"]), f"{tag} {nm}"


@pytest.mark.parametrize("fused", [False, True], ids=["torch.optim.Adam", "FusedAdam"])
def test_prune_points_matches_reference(fused):
    from lidar_rt_b200 import densify
    from lidar_rt_b200.optim import FusedAdam
    g = load_golden("ref_densify.npz")
    m = Model(g, FusedAdam if fused else torch.optim.Adam)
    densify.prune_points(m, torch.tensor(g["prune/mask"], device="cuda"))
    _compare(m, g, "prune")


@pytest.mark.parametrize("fused", [False, True], ids=["torch.optim.Adam", "FusedAdam"])
def test_densify_and_prune_matches_reference(fused):
    from lidar_rt_b200 import densify
    from lidar_rt_b200.optim import FusedAdam
    g = load_golden("ref_densify.npz")
    m = Model(g, FusedAdam if fused else torch.optim.Adam)
    o = g["opt"]
    opt = types.SimpleNamespace(densify_grad_threshold=float(o[0]), thresh_opa_prune=float(o[1]), prune_size_threshold=float(o[2]))
    res = densify.densify_and_prune(m, opt, 0.005, 20, samples=torch.tensor(g["densify/samples"], device="cuda"))
    assert tuple(int(x) for x in res) == tuple(int(x) for x in g["densify/counts"])
    # rows that are plain moves: everything except the split children that survived the final prune; compare those by tolerance
    clone_n, split_n = int(res[0]), int(res[1])
    P = g["in/xyz"].shape[0]
    # survivors of the final prune among the first (P - split_n + clone_n) rows come first and are exact copies
    low = 1.0 / (1.0 + np.exp(-g["densify/opacity"][:, 0]))
    n_exact = 0                       # conservative: count leading rows whose scaling is bit-identical to some input row
    inp = {tuple(r) for r in g["in/scaling"].tolist()}
    for r in g["densify/scaling"].tolist():
        if tuple(r) in inp:
            n_exact += 1
        else:
            break
    assert n_exact > 0 and low.min() >= float(o[1]) - 1e-6
    _compare(m, g, "densify", exact_rows=n_exact)
    # the restructured model trains on: one optimiser step with both optimisers' state intact
    for n, a in NAMES:
        getattr(m, a).grad = torch.randn_like(getattr(m, a))
    m.optimizer.step()
    assert all(torch.isfinite(getattr(m, a)).all() for _, a in NAMES)


def test_compact_rows_edge_cases():
    from lidar_rt_b200 import native
    ctx = native.Context()
    src = torch.arange(0, 70, dtype=torch.float32, device="cuda").reshape(10, 7)
    for keep in (torch.ones(10, dtype=torch.bool, device="cuda"), torch.zeros(10, dtype=torch.bool, device="cuda"),
                 torch.tensor([1, 0, 0, 1, 1, 0, 1, 0, 0, 1], dtype=torch.bool, device="cuda")):
        dst = torch.full((int(keep.sum()), 7), -1.0, device="cuda")
        vec = torch.arange(10, dtype=torch.float32, device="cuda"); dvec = torch.full((int(keep.sum()),), -1.0, device="cuda")
        if dst.shape[0] == 0:
            dst = torch.empty((0, 7), device="cuda"); dvec = torch.empty((0,), device="cuda")
        ctx.compact_rows(keep, [(src, dst, 0), (vec, dvec, 0)])
        assert torch.equal(dst, src[keep]) and torch.equal(dvec, vec[keep])
    with pytest.raises(native.LrtError):
        ctx.compact_rows(torch.ones(10, dtype=torch.bool, device="cuda"), [(src, torch.empty((10, 6), device="cuda"), 0)])
    ctx.close()

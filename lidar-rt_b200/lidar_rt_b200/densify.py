"""Densify / prune of a GaussianModel through the native row-compaction kernels (second half).

Host-side mirror of the reference's restructuring methods (/root/reference/lib/scene/gaussian_model.py):
    prune_points(model, mask)                                   :253-270 (+ _prune_optimizer :235-251)
    densify_and_clone_split(model, grads, grad_threshold, N=2)  densify_and_clone :338-352 followed by densify_and_split :311-336
    densify_and_prune(model, opt, min_opacity, max_screen_size) :354-407
over a model that is duck-typed like the reference's GaussianModel: the six leaf parameters `_xyz, _features_dc, _features_rest,
_opacity, _scaling, _rotation`, an `optimizer` with one single-parameter group per leaf named like the reference's
(training_setup :192-199: xyz, f_dc, f_rest, opacity, scaling, rotation; torch.optim.Adam or optim.FusedAdam), the statistics
`xyz_gradient_accum, denom, max_radii2D`, and `densify_scale_threshold, extent, get_scaling, get_opacity`.

Same results as the reference's torch ops — the same rows in the same order, parameters and Adam moments moved bit for bit,
new moment rows zero, statistics reset as densification_postfix does (:304-306) — but every tensor that has a row per Gaussian
goes through ONE native call per restructuring step instead of ~40 torch indexing / cat kernels that each re-allocate.
The normal samples of the split are drawn here with torch.normal exactly like the reference (same generator consumption).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import native

GROUPS = (("xyz", "_xyz", native.ROW_XYZ), ("f_dc", "_features_dc", native.ROW_COPY), ("f_rest", "_features_rest", native.ROW_COPY),
          ("opacity", "_opacity", native.ROW_COPY), ("scaling", "_scaling", native.ROW_SCALING), ("rotation", "_rotation", native.ROW_COPY))

_ctx: dict = {}


def _context(device) -> native.Context:
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _ctx:
        _ctx[key] = native.Context(torch.device("cuda", key))
    return _ctx[key]


def _group(model, name):
    for g in model.optimizer.param_groups:
        if g["name"] == name:
            assert len(g["params"]) == 1
            return g
    raise KeyError(name)


def _rows(model, n_out):
    """(src, dst, kind) triples of every per-Gaussian tensor: parameters and, where the optimizer has state, both moments."""
    rows, new = [], {}
    for name, attr, kind in GROUPS:
        g = _group(model, name)
        p = g["params"][0]
        src = p.detach()
        if not src.is_contiguous():
            src = src.contiguous()
        dst = torch.empty((n_out,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        rows.append((src, dst, kind))
        st = model.optimizer.state.get(p, None)
        moments = None
        if st is not None and "exp_avg" in st:
            moments = []
            for k in ("exp_avg", "exp_avg_sq"):
                m = st[k].contiguous()
                d = torch.empty((n_out,) + tuple(m.shape[1:]), dtype=m.dtype, device=m.device)
                rows.append((m, d, native.ROW_ZERO_NEW))
                moments.append(d)
        new[name] = (g, p, st, dst, moments)
    return rows, new


def _install(model, new):
    """What _prune_optimizer / cat_tensors_to_optimizer do with the results (:244-248, :284-288): a fresh nn.Parameter per group,
    the state dict re-keyed to it with the new moments (its `step` untouched)."""
    for name, attr, _ in GROUPS:
        g, p_old, st, dst, moments = new[name]
        p_new = nn.Parameter(dst.requires_grad_(True))
        if st is not None:
            if moments is not None:
                st["exp_avg"], st["exp_avg_sq"] = moments
            del model.optimizer.state[p_old]
            model.optimizer.state[p_new] = st
        g["params"][0] = p_new
        setattr(model, attr, p_new)


def prune_points(model, mask: torch.Tensor, ctx: Optional[native.Context] = None):
    """gaussian_model.py:253-270: drop the rows where `mask` is True."""
    keep = ~mask.reshape(-1).bool()
    n_keep = int(keep.sum().item())
    ctx = ctx or _context(model._xyz.device)
    rows, new = _rows(model, n_keep)
    stats = []
    for nm in ("xyz_gradient_accum", "denom", "max_radii2D"):
        t = getattr(model, nm, None)
        if isinstance(t, torch.Tensor) and t.dim() >= 1 and t.shape[0] == keep.shape[0]:
            d = torch.empty((n_keep,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            rows.append((t.contiguous(), d, native.ROW_COPY))
            stats.append((nm, d))
    ctx.compact_rows(keep, rows)
    _install(model, new)
    for nm, d in stats:
        setattr(model, nm, d)


def densify_and_clone_split(model, grads: torch.Tensor, grad_threshold: float, N: int = 2, ctx: Optional[native.Context] = None,
                            samples: Optional[torch.Tensor] = None):
    """densify_and_clone (:338-352) followed by densify_and_split (:311-336), as one restructuring pass.
    Returns (clone_num, split_num) like the two calls in densify_and_prune (:357-358)."""
    P = model._xyz.shape[0]
    grads = grads.reshape(-1)
    scal_max = torch.max(model.get_scaling, dim=1).values
    thr = model.densify_scale_threshold * model.extent
    grad_mask = grads >= grad_threshold
    clone_mask = torch.logical_and(grad_mask, scal_max <= thr)               # :340-342
    # the split sees the P + n_clone rows after the clone step with `padded` gradients: zeros for the clones (:314-316), and the
    # clones inherit a scale <= thr, so only original rows can be selected
    split_mask = torch.logical_and(grad_mask, scal_max > thr)                # :317-318
    n_clone, n_split = int(clone_mask.sum().item()), int(split_mask.sum().item())
    if samples is None:                                                      # :321-325, same calls -> same generator consumption
        stds = model.get_scaling[split_mask].repeat(N, 1)
        if getattr(model, "dimension", 2) == 2:
            stds = torch.cat([stds, 0 * torch.ones_like(stds[:, :1])], dim=-1)
        samples = torch.normal(mean=torch.zeros_like(stds), std=stds)
    n_out = P - n_split + n_clone + N * n_split
    ctx = ctx or _context(model._xyz.device)
    rows, new = _rows(model, n_out)
    ctx.densify_rows(clone_mask, split_mask, n_clone, n_split, N, samples, model._rotation.detach(), rows)
    _install(model, new)
    dev = model._xyz.device
    model.xyz_gradient_accum = torch.zeros((n_out, 1), device=dev)           # densification_postfix :304-306, then pruned: still zeros
    model.denom = torch.zeros((n_out, 1), device=dev)
    model.max_radii2D = torch.zeros((n_out,), device=dev)
    return n_clone, n_split


def densify_and_prune(model, opt, min_opacity, max_screen_size, ctx: Optional[native.Context] = None, samples: Optional[torch.Tensor] = None):
    """gaussian_model.py:354-407. `samples`: the split's normal samples if the caller drew them already (tests do)."""
    mean_grads = (model.xyz_gradient_accum / model.denom).nan_to_num(0.0).squeeze(-1)
    clone_num, split_num = densify_and_clone_split(model, mean_grads, opt.densify_grad_threshold, ctx=ctx, samples=samples)
    low_opacity = (model.get_opacity < opt.thresh_opa_prune).squeeze()
    prune_mask = low_opacity
    prune_opacity_num = int(low_opacity.sum().item())
    prune_scale_num = 0
    if max_screen_size:
        big_points_ws = model.get_scaling.max(dim=1).values > 0.1 * model.extent * opt.prune_size_threshold
        prune_scale_num = int(big_points_ws.sum().item())
        prune_mask = torch.logical_or(low_opacity, big_points_ws)
        bb = getattr(model, "bounding_box", None)
        if bb is not None and hasattr(bb, "min_xyz"):                        # actors: samples outside the tracking box (:377-400)
            from .densify_ref_ops import points_outside_box
            prune_mask = torch.logical_or(prune_mask, points_outside_box(model, bb))
    if prune_mask.sum() < model._xyz.shape[0]:
        prune_points(model, prune_mask, ctx=ctx)
    return clone_num, split_num, prune_scale_num, prune_opacity_num

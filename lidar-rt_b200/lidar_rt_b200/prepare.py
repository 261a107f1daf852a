"""Fused parameter activation + world transform + per-asset concatenation.

`fused_prepare(gaussian_assets, frame, dynamic, decomp)` returns the five tensors the reference's render glue
assembles with ~20 torch kernels (/root/reference/lib/gaussian_renderer/__init__.py:76-134 over the accessors of
/root/reference/lib/scene/gaussian_model.py:112-148) — means3D, opacity, scales, rotations, shs — from ONE native
call that reads the GaussianModel leaves (`_xyz`, `_scaling`, `_rotation`, `_opacity`, `_features_dc`,
`_features_rest`) in place, and whose backward writes the leaf gradients in one call as well.

It applies when every asset exposes those leaves; anything else (duck-typed assets with only accessors,
precomputed covariances / colours) returns None and the caller keeps the accessor path.
"""
from __future__ import annotations

import torch

from . import native

_LEAVES = ("_xyz", "_scaling", "_rotation", "_opacity", "_features_dc", "_features_rest")
_NAMES = ("xyz", "scaling", "rotation", "opacity", "features_dc", "features_rest")


def _pose(pc, frame):
    """(T (3,), quat (4,)) of the asset at `frame`, or None: BoundingBox.frame[t][:2] in the reference
    (lib/scene/bounding_box.py:53), `actor_poses[t]` in this repo's stand-in."""
    bb = getattr(pc, "bounding_box", None)
    fr = getattr(bb, "frame", None) if bb is not None else getattr(pc, "actor_poses", None)
    if fr is not None and frame in fr:
        return fr[frame][0], fr[frame][1]
    return None


class _FusedPrepare(torch.autograd.Function):
    @staticmethod
    def forward(ctx, nctx, specs, *leaves):
        assets = []
        for k, sp in enumerate(specs):
            a = {nm: leaves[6 * k + i].detach() for i, nm in enumerate(_NAMES)}
            a.update(sp)
            assets.append(a)
        out = nctx.prepare(assets)
        ctx.nctx, ctx.assets = nctx, assets
        ctx.want = [{nm: ctx.needs_input_grad[2 + 6 * k + i] for i, nm in enumerate(_NAMES)} for k in range(len(specs))]
        return out

    @staticmethod
    def backward(ctx, g_means, g_scales, g_rots, g_opac, g_shs):
        z = lambda g, ref_shape: torch.zeros(ref_shape, device=ctx.nctx.device) if g is None else g.contiguous()
        P = sum(a["xyz"].shape[0] for a in ctx.assets)
        M = 1 + ctx.assets[0]["features_rest"].shape[1]
        grads = ctx.nctx.prepare_backward(ctx.assets, z(g_means, (P, 3)), z(g_scales, (P, 2)), z(g_rots, (P, 4)), z(g_opac, (P, 1)),
                                          z(g_shs, (P, M, 3)), want=ctx.want)
        flat = []
        for g in grads:
            flat += [g[nm] for nm in _NAMES]
        return (None, None, *flat)


def fused_prepare(gaussian_assets, frame, dynamic: bool, decomp, nctx: native.Context):
    """-> (means3D, opacity, scales, rotations, shs) or None if the fused path does not apply."""
    if not gaussian_assets or len(gaussian_assets) > native.MAX_ASSETS:
        return None
    specs, leaves = [], []
    for k, pc in enumerate(gaussian_assets):
        ts = [getattr(pc, nm, None) for nm in _LEAVES]
        if any(not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.dim() == 0
               for t in ts):
            return None
        if ts[4].dim() != 3 or ts[4].shape[1] != 1 or ts[5].dim() != 3:
            return None
        # which assets get the composed rotation: reference :117-130
        if decomp == "background" or not dynamic:
            if k > 0:
                return None           # the reference takes rot_in_local[0] only (:118): a static scene has one asset
            compose = False
        elif decomp == "object":
            compose = True
        else:
            compose = k > 0
        sp = {"compose_rotation": compose}
        pose = _pose(pc, frame)
        if pose is not None:
            T, q = pose
            if not (isinstance(T, torch.Tensor) and isinstance(q, torch.Tensor) and T.is_cuda and q.is_cuda):
                return None
            sp["pose_T"] = T.detach().reshape(3).to(torch.float32).contiguous()
            sp["pose_quat"] = q.detach().reshape(4).to(torch.float32).contiguous()
        elif compose:
            return None           # an actor without a pose at this frame: the reference yields a zero quaternion; keep its path
        specs.append(sp)
        leaves += ts
    means, scales, rots, opac, shs = _FusedPrepare.apply(nctx, specs, *leaves)
    return means, opac, scales, rots, shs

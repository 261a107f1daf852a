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


class ShInPlace:
    """What travels with the SH handle tensor when the coefficients are NOT concatenated (fused_prepare(sh_in_place=True)): the
    leaves the tracer reads in place (lrt_set_sh_parts) and, after the tracer's backward, the leaf gradients it wrote in place."""

    def __init__(self, parts, want):
        self.parts = parts            # [(features_dc, features_rest)] detached, concatenation order
        self.want = want              # [(bool, bool)]: which leaf gradients are needed
        self.grads = None             # [(d_features_dc | None, d_features_rest | None)], set by _Tracer.backward
        self.M = 1 + parts[0][1].shape[1]
        self.P = sum(dc.shape[0] for dc, _ in parts)

    def gradient_buffers(self, device):
        """One allocation per leaf kind; every asset's `rest` block starts on a 16-byte boundary (vector reductions)."""
        n_dc = sum(dc.numel() for dc, _ in self.parts)
        offs, n_rest = [], 0
        for _, rest in self.parts:
            offs.append(n_rest)
            n_rest += (rest.numel() + 3) & ~3
        f_dc = torch.empty(n_dc, dtype=torch.float32, device=device); f_rest = torch.empty(n_rest, dtype=torch.float32, device=device)
        out, o = [], 0
        for k, (dc, rest) in enumerate(self.parts):
            gdc = f_dc[o:o + dc.numel()].view(dc.shape) if self.want[k][0] else None
            grest = f_rest[offs[k]:offs[k] + rest.numel()].view(rest.shape) if self.want[k][1] else None
            o += dc.numel()
            out.append((gdc, grest))
        return out


class _FusedPrepare(torch.autograd.Function):
    @staticmethod
    def forward(ctx, nctx, specs, box, *leaves):
        assets = []
        for k, sp in enumerate(specs):
            a = {nm: leaves[6 * k + i].detach() for i, nm in enumerate(_NAMES)}
            a.update(sp)
            assets.append(a)
        means, scales, rots, opac, shs = nctx.prepare(assets, with_shs=box is None)
        ctx.nctx, ctx.assets, ctx.box = nctx, assets, box
        ctx.want = [{nm: ctx.needs_input_grad[3 + 6 * k + i] for i, nm in enumerate(_NAMES)} for k in range(len(specs))]
        if box is not None:
            # the SH handle: shaped like the concatenated tensor, backed by ONE float (stride 0). It is only ever handed to the
            # tracer, which reads the leaves through `box` instead.
            box.want = [(w["features_dc"], w["features_rest"]) for w in ctx.want]
            shs = torch.zeros(1, dtype=torch.float32, device=means.device).expand(means.shape[0], box.M, 3)
        return means, scales, rots, opac, shs

    @staticmethod
    def backward(ctx, g_means, g_scales, g_rots, g_opac, g_shs):
        z = lambda g, ref_shape: torch.zeros(ref_shape, device=ctx.nctx.device) if g is None else g.contiguous()
        P = sum(a["xyz"].shape[0] for a in ctx.assets)
        M = 1 + ctx.assets[0]["features_rest"].shape[1]
        box = ctx.box
        grads = ctx.nctx.prepare_backward(ctx.assets, z(g_means, (P, 3)), z(g_scales, (P, 2)), z(g_rots, (P, 4)), z(g_opac, (P, 1)),
                                          None if box is not None else z(g_shs, (P, M, 3)), want=ctx.want)
        if box is not None:            # the tracer's backward wrote the SH leaf gradients in place (or never ran: no gradient)
            for k, g in enumerate(grads):
                gdc, grest = box.grads[k] if box.grads is not None else (None, None)
                g["features_dc"], g["features_rest"] = gdc, grest
            box.grads = None
        flat = []
        for g in grads:
            flat += [g[nm] for nm in _NAMES]
        return (None, None, None, *flat)


def fused_prepare(gaussian_assets, frame, dynamic: bool, decomp, nctx: native.Context, sh_in_place: bool = False):
    """-> (means3D, opacity, scales, rotations, shs) or None if the fused path does not apply.
    sh_in_place=True: `shs` is only a HANDLE for diff_lidar_tracer.Tracer (attribute `_lrt_sh`): no concatenated copy of the SH
    coefficients is made, the tracer reads features_dc / features_rest in place and its backward writes their gradients in place."""
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
    box = None
    if sh_in_place:
        box = ShInPlace([(leaves[6 * k + 4].detach(), leaves[6 * k + 5].detach()) for k in range(len(specs))], None)
    means, scales, rots, opac, shs = _FusedPrepare.apply(nctx, specs, box, *leaves)
    if box is not None:
        shs._lrt_sh = box
    return means, opac, scales, rots, shs

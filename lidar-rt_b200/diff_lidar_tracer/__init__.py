"""Drop-in replacement for the reference package `diff_lidar_tracer`
(/root/reference/submodules/diff-lidar-tracer/diff_lidar_tracer/__init__.py:13-219).

Same public names, argument order and return shapes:
    Tracer()                       nn.Module, no arguments, safe to construct at import time
    Tracer.build_acceleration_structure(vertices, triangles, rebuild=True)
    Tracer.forward(ray_o, ray_d, mesh_normals, means3D, grads3D, shs, colors_precomp, opacities,
                   scales, rotations, cov3Ds_precomp, tracer_settings)
        -> (rendered (H, W, 9) float32, accum_gaussian_weights (P,) float32)
    TracingSettings                NamedTuple with the reference's 12 fields
    _Tracer                        torch.autograd.Function, 14 inputs, backward returns 14 entries

What is different underneath: no OptiX. The acceleration structure is an LBVH over analytic proxy
quads derived from (means3D, scales, rotations, opacities) — the same quads build2DRectangle
tessellates — so `vertices` / `triangles` are accepted for compatibility but not read; the
structure is (re)built lazily by the next forward() from the Gaussian parameters.
`Tracer.build_from_gaussians()` builds it directly and lets callers skip build2DRectangle.
"""
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from lidar_rt_b200 import native

__all__ = ["Tracer", "TracingSettings", "_Tracer"]


class TracingSettings(NamedTuple):          # reference :139-151 (only bg, scale_modifier, sh_degree are live)
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class _Handle:
    """What the reference passes around as `optix_context` (:159-162): the native context plus the
    build request recorded by build_acceleration_structure()."""

    def __init__(self):
        native.load_library()           # fail at construction if the .so was never built
        self._ctx: Optional[native.Context] = None
        self.pending: Optional[str] = "build"
        self.hit_cap = native.DEFAULT_HIT_CAP
        self.record_hits = True
        self.bwd_flags = 0
        # The reference reads means3D / scales / rotations / opacities on EVERY forward (forward.cu:228-251) and uses the
        # mesh only for the BVH, so a caller may build once and keep tracing while the parameters change. Here the
        # structure holds derived surfel records, so a forward without a pending build request refits them first
        # (k_records + k_fit, ~0.17 ms at 2 M Gaussians). Tensor identity cannot stand in for "unchanged": activations
        # return fresh tensors that the caching allocator places at the same address. Callers that KNOW the parameters
        # are those of the last build (a sweep over static Gaussians) may set this to True and skip the refit.
        self.assume_static = False

    @property
    def ctx(self) -> native.Context:
        if self._ctx is None:
            self._ctx = native.Context()
        return self._ctx

    def ensure_built(self, means3D, scales, rotations, opacities, scale_modifier):
        ctx = self.ctx
        info_P = ctx.info().P
        P = means3D.shape[0]
        if self.pending == "build" or info_P != P:
            ctx.build(means3D, scales, rotations, opacities, scale_modifier, refit=False)
        elif self.pending == "refit" or not self.assume_static:
            ctx.build(means3D, scales, rotations, opacities, scale_modifier, refit=True)
        self.pending = None
        return ctx.generation


def _is_empty(t) -> bool:
    return t is None or (isinstance(t, torch.Tensor) and t.numel() == 0)


class _Tracer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, handle, training, ray_o, ray_d, vertices, means3D, grads3D, shs, colors_precomp, opacities,
                scales, rotations, cov3Ds_precomp, tracer_settings):
        if _is_empty(shs) or not _is_empty(colors_precomp):
            raise NotImplementedError("only the SH colour path is implemented natively (as in the reference device code)")
        if _is_empty(scales) or _is_empty(rotations) or not _is_empty(cov3Ds_precomp):
            raise NotImplementedError("only the scale/rotation path is implemented natively (as in the reference device code)")
        s = tracer_settings
        means_d, scales_d, rots_d, opac_d, shs_d = (t.detach() for t in (means3D, scales, rotations, opacities, shs))
        gen = handle.ensure_built(means_d, scales_d, rots_d, opac_d, float(s.scale_modifier))
        # SH handle of lidar_rt_b200.prepare.fused_prepare(sh_in_place=True): no concatenated coefficients exist; the library
        # reads the model's features_dc / features_rest in place (lrt_set_sh_parts)
        box = getattr(shs, "_lrt_sh", None)
        ctx.sh_box = box
        sh_M = 16
        if box is not None:
            if box.P != means3D.shape[0]:
                raise ValueError("SH handle and means3D disagree on the number of Gaussians")
            handle.ctx.set_sh_parts(box.parts)
            shs_d, sh_M = None, box.M
        # hit lists are recorded for the backward's replay only: not under torch.no_grad() / when no input needs a gradient
        # (the forward then keeps its lists in the context's own workspace instead of ~1 GB of fresh tensors per call)
        record = handle.record_hits and any(ctx.needs_input_grad)
        res = handle.ctx.forward(ray_o.detach(), ray_d.detach(), s.bg, means_d, scales_d, rots_d, opac_d, shs_d,
                                 int(s.sh_degree), float(s.scale_modifier), record_hits=record,
                                 cap=handle.hit_cap, sh_M=sh_M)
        out = res["out"]
        accum = res["accum_w"]
        ctx.handle = handle
        ctx.settings = s
        ctx.generation = gen
        ctx.cap = res["cap"]
        ctx.have_hits = res["hit_gidx"] is not None
        saved = [ray_o, ray_d, means3D, shs, opacities, scales, rotations, out]
        if ctx.have_hits:
            saved += [res["hit_gidx"], res["hit_t"], res["hit_cnt"], res["hit_aux"]]
        ctx.save_for_backward(*saved)
        ctx.mark_non_differentiable(accum)
        return out, accum

    @staticmethod
    def backward(ctx, grad_out, _grad_accum):
        handle, s = ctx.handle, ctx.settings
        saved = ctx.saved_tensors
        ray_o, ray_d, means3D, shs, opacities, scales, rotations, out = saved[:8]
        hits = None
        if ctx.have_hits:
            hits = dict(hit_gidx=saved[8], hit_t=saved[9], hit_cnt=saved[10], hit_aux=saved[11], cap=ctx.cap)
        nctx = handle.ctx
        stale = nctx.generation != ctx.generation
        if stale and (hits is None or bool((hits["hit_cnt"] > ctx.cap).any())):
            # rays that must be re-traced need the structure of THIS forward: rebuild it
            nctx.build(means3D.detach(), scales.detach(), rotations.detach(), opacities.detach(),
                       float(s.scale_modifier))
            handle.pending = "build"          # whatever was current is gone
        box = ctx.sh_box
        if box is not None:              # SH gradients straight into the leaf gradients; the handle's own gradient carries nothing
            box.grads = box.gradient_buffers(means3D.device)
            nctx.set_sh_parts(box.parts, box.grads)
        g = nctx.backward(ray_o, ray_d, s.bg, means3D.detach(), scales.detach(), rotations.detach(),
                          opacities.detach(), None if box is not None else shs.detach(), int(s.sh_degree), out, grad_out.contiguous(),
                          hits=hits, scale_modifier=float(s.scale_modifier), flags=handle.bwd_flags, sh_M=box.M if box is not None else 16)
        if box is not None:
            g["shs"] = torch.zeros(1, dtype=torch.float32, device=means3D.device).expand(shs.shape)
        grad_opac = g["opac"].reshape(opacities.shape)
        zeros3 = torch.zeros_like(means3D)      # reference returns zero grads for the unused slots (:322-329)
        return (None, None, None, None, None,
                g["means"], zeros3, g["shs"], None, grad_opac, g["scales"], g["rots"], None, None)


class Tracer(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.optix_context = _Handle()          # attribute name kept from the reference (:162)
        self.vertices = None

    # -- reference API -------------------------------------------------------------------------
    def build_acceleration_structure(self, vertices: torch.Tensor, triangles: torch.Tensor, rebuild: bool = True):
        """Reference :164-171. Records the request; the LBVH is built from the Gaussian parameters
        on the next forward() (rebuild=True -> full build, False -> refit in the stored order)."""
        self.vertices = vertices
        self.optix_context.pending = "build" if rebuild else "refit"

    def forward(self, ray_o, ray_d, mesh_normals, means3D, grads3D, shs=None, colors_precomp=None, opacities=None,
                scales=None, rotations=None, cov3Ds_precomp=None, tracer_settings: TracingSettings = None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3Ds_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3Ds_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        return _Tracer.apply(self.optix_context, self.training, ray_o, ray_d, self.vertices, means3D, grads3D,
                             shs, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, tracer_settings)

    # -- extensions ----------------------------------------------------------------------------
    def build_from_gaussians(self, means3D, scales, rotations, opacities, scale_modifier: float = 1.0,
                             rebuild: bool = True):
        """Build (or refit) the LBVH now, straight from the Gaussian parameters; replaces
        build2DRectangle + build_acceleration_structure."""
        h = self.optix_context
        h.ctx.build(means3D.detach(), scales.detach(), rotations.detach(), opacities.detach(), scale_modifier,
                    refit=not rebuild and h.ctx.info().P == means3D.shape[0])
        h.pending = None

    def info(self):
        return self.optix_context.ctx.info()

"""One-launch Adam for Gaussian assets, and the packed gradient all-reduce of data-parallel training.

`FusedAdam` is a drop-in for the `torch.optim.Adam(l, lr=0.0, eps=1e-15)` every GaussianModel of the reference creates
(/root/reference/lib/scene/gaussian_model.py:186-201): same constructor arguments, same `param_groups` (the reference's
`update_learning_rate` writes `group['lr']`, :207-213) and the same per-parameter state — `state[p] = {"step", "exp_avg",
"exp_avg_sq"}` — which the reference's densification code edits directly (`replace_tensor_to_optimizer`,
`_prune_optimizer`, `cat_tensors_to_optimizer`, :220-298). `step()` hands every tensor to ONE native call
(`lrt_adam_step`); `step_many(optimizers)` does the same for the optimizers of all assets of a scene together — the
reference steps them one after the other, ~2000 small launches per iteration on a 41-asset scene.

`all_reduce_gradients(params)` is the collective of data-parallel training over frames (each rank renders and
back-propagates different frames of the same Gaussians): the gradients of all parameters travel as ONE packed buffer through
one all-reduce (NCCL on GPUs; any torch.distributed backend works) and come back as views into it.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import native

_ctx: dict = {}


def _context(device: torch.device) -> native.Context:
    key = device.index if device.index is not None else torch.cuda.current_device()
    c = _ctx.get(key)
    if c is None:
        c = _ctx[key] = native.Context(torch.device("cuda", key))
    return c


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam's interface and state layout (amsgrad / weight decay / maximize are not offered: the reference uses none)."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    def _rows(self):
        """[(betas, eps, [(param, grad, exp_avg, exp_avg_sq, lr, step), ...])]; creates state lazily and advances `step`."""
        out = {}
        for group in self.param_groups:
            key = (float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]))
            rows = out.setdefault(key, [])
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] = st["step"] + 1 if isinstance(st["step"], torch.Tensor) else torch.tensor(float(st["step"]) + 1)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                rows.append((p.data, g, st["exp_avg"], st["exp_avg_sq"], float(group["lr"]), int(st["step"])))
        return out

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        _launch([self])
        return loss


def _launch(optimizers: Sequence[FusedAdam]):
    merged = {}
    for o in optimizers:
        for key, rows in o._rows().items():
            merged.setdefault(key, []).extend(rows)
    for (b1, b2, eps), rows in merged.items():
        by_dev = {}
        for r in rows:
            by_dev.setdefault(r[0].device, []).append(r)
        for dev, rs in by_dev.items():
            if dev.type != "cuda":
                raise native.LrtError("FusedAdam needs CUDA parameters; there is no CPU fallback")
            _context(dev).adam_step(rs, b1, b2, eps)


@torch.no_grad()
def step_many(optimizers: Iterable[FusedAdam]):
    """One native call for the optimizers of all assets (they share betas / eps in the reference, so: one launch)."""
    _launch(list(optimizers))


@torch.no_grad()
def all_reduce_gradients(params: Iterable[torch.Tensor], world_size: Optional[int] = None, group=None,
                         average: bool = True) -> Optional[torch.Tensor]:
    """Sum (or average) the gradients of `params` over the ranks with ONE collective over a packed buffer; afterwards every
    `p.grad` is a view into that buffer (which is returned). Parameters without a gradient on this rank count as zero —
    every rank must pass the same parameters in the same order."""
    ps: List[torch.Tensor] = [p for p in params]
    if not ps:
        return None
    world = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
    total = sum(p.numel() for p in ps)
    flat = torch.empty(total, dtype=ps[0].dtype, device=ps[0].device)
    off = 0
    views = []
    for p in ps:
        v = flat[off:off + p.numel()].view_as(p)
        if p.grad is None:
            v.zero_()
        else:
            v.copy_(p.grad)
        views.append(v)
        off += p.numel()
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(world)
    for p, v in zip(ps, views):
        p.grad = v
    return flat

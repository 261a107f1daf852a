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

import numpy as np
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


_ROW = np.dtype([("param", np.uint64), ("grad", np.uint64), ("exp_avg", np.uint64), ("exp_avg_sq", np.uint64),
                 ("n", np.int64), ("lr", np.float32), ("step", np.int32)])          # lrt_adam_tensor, 48 bytes


class _State(dict):
    """Per-parameter state dict that tells its optimizer when somebody replaces an entry (the reference's densification code
    assigns new exp_avg / exp_avg_sq tensors into it, gaussian_model.py:224-231)."""
    __slots__ = ("owner",)

    def __setitem__(self, k, v):
        o = getattr(self, "owner", None)
        if o is not None and not o._dirty:
            o._flush_steps()          # bring "step" up to date before the first outside edit, so an edit of "step" itself sticks
            o._dirty = True
        dict.__setitem__(self, k, v)


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam's interface and state layout (amsgrad / weight decay / maximize are not offered: the reference uses none).

    The table handed to `lrt_adam_step` is kept between steps (a numpy array laid out like `lrt_adam_tensor`): as long as the
    parameters and their state tensors are the same objects, a step only refreshes the gradient pointers, learning rates and
    step counts. `state[p]["step"]` is brought up to date whenever the table is rebuilt and by `state_dict()`."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._dirty = True
        self._key = None

    # ---- table
    def _flush_steps(self):
        """write the cached step counts back into state[p]["step"]"""
        if self._key is None:
            return
        for p, n in zip(self._ps, self._tab["step"]):
            st = self.state.get(p)
            if st is not None and "step" in st:
                dict.__setitem__(st, "step", torch.tensor(float(n)))

    def _rebuild(self, ps, groups_of):
        if not self._dirty:
            self._flush_steps()       # the parameter set changed: the counts of the old table go back into the state first
        tab = np.zeros(len(ps), _ROW)
        for k, p in enumerate(ps):
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise native.LrtError("FusedAdam needs contiguous float32 CUDA parameters; there is no CPU fallback")
            st = self.state[p]
            if type(st) is not _State:
                st = _State(st); self.state[p] = st
            st.owner = None
            if "exp_avg" not in st:
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            m, v = st["exp_avg"], st["exp_avg_sq"]
            for t, nm in ((m, "exp_avg"), (v, "exp_avg_sq")):
                if t.device != p.device or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != p.numel():
                    raise native.LrtError(f"FusedAdam: state {nm} does not match its parameter")
            st.owner = self
            tab[k] = (p.data_ptr(), 0, m.data_ptr(), v.data_ptr(), p.numel(), 0.0, int(st["step"]))
        self._ps, self._tab, self._groups_of = ps, tab, np.asarray(groups_of, np.int64)
        self._key = tuple(map(id, ps))
        self._dirty = False

    def _prepare(self):
        """-> (table rows of the parameters that have a gradient) with grad / lr / step refreshed"""
        ps, groups_of = [], []
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                if p.grad is not None:
                    ps.append(p); groups_of.append(gi)
        if not ps:
            return None
        if self._dirty or self._key != tuple(map(id, ps)):
            self._rebuild(ps, groups_of)
        tab = self._tab
        grads = [p.grad for p in ps]
        for g, p in zip(grads, ps):
            if g.dtype != torch.float32 or g.is_sparse or g.device != p.device or g.numel() != p.numel():
                raise native.LrtError("FusedAdam: gradients must be dense float32 tensors on the parameter's device")
        grads = [g if g.is_contiguous() else g.contiguous() for g in grads]
        tab["grad"] = [g.data_ptr() for g in grads]
        tab["lr"] = np.asarray([float(g["lr"]) for g in self.param_groups], np.float32)[self._groups_of]
        tab["step"] += 1
        return tab, grads          # grads kept alive by the caller until the launch is enqueued

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        _launch([self])
        return loss

    def state_dict(self):
        self._flush_steps()
        sd = super().state_dict()
        sd["state"] = {k: dict(v) for k, v in sd["state"].items()}      # plain dicts: nothing of this optimizer travels with a saved state
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._dirty = True
        self._key = None


def _launch(optimizers: Sequence[FusedAdam]):
    buckets = {}
    keep = []
    for o in optimizers:
        hp = {(float(g["betas"][0]), float(g["betas"][1]), float(g["eps"])) for g in o.param_groups}
        if len(hp) != 1:
            raise native.LrtError("FusedAdam: all parameter groups of an optimizer must share betas and eps (the reference's do)")
        r = o._prepare()
        if r is None:
            continue
        tab, grads = r
        keep.append(grads)
        dev = o._ps[0].device
        if any(p.device != dev for p in o._ps):
            raise native.LrtError("FusedAdam: the parameters of an optimizer must live on one device")
        buckets.setdefault((dev, next(iter(hp))), []).append(tab)
    for (dev, (b1, b2, eps)), tabs in buckets.items():
        _context(dev).adam_step_table(tabs[0] if len(tabs) == 1 else np.concatenate(tabs), b1, b2, eps)


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

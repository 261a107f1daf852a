"""Frame-parallel multi-GPU sweeps.

The reference is single-process / single-GPU; every `raytracing()` call is independent given the
Gaussian set (/root/reference/lib/gaussian_renderer/__init__.py:142-147 rebuilds the acceleration
structure per call), so a sweep shards by FRAME: one process per GPU (torchrun), Gaussians
replicated, rank r renders frames f = r (mod world). There is no collective on the data path;
NCCL is used once per sweep to gather the rendered (H, W, 9) buffers.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch
import torch.distributed as dist


def frames_of_rank(n_frames: int, rank: int, world: int) -> List[int]:
    """Round-robin shard: frame f belongs to rank f % world."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    return list(range(rank, n_frames, world))


def gather_frames(local: torch.Tensor, n_frames: int, rank: int, world: int) -> torch.Tensor:
    """all_gather the per-rank stacks (n_local, ...) into frame order (n_frames, ...).

    Ranks may hold different counts (n_frames % world != 0): stacks are padded to the maximum so a
    single all_gather_into_tensor suffices (NCCL) — or all_gather for backends without it (gloo)."""
    if world == 1:
        return local
    n_max = (n_frames + world - 1) // world
    pad = n_max - local.shape[0]
    if pad:
        local = torch.cat([local, local.new_zeros((pad,) + tuple(local.shape[1:]))], 0)
    local = local.contiguous()
    if dist.get_backend() == "nccl":
        flat = local.new_empty((world * n_max,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(flat, local)
        parts = flat.view((world, n_max) + tuple(local.shape[1:]))
    else:
        lst = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(lst, local)
        parts = torch.stack(lst, 0)
    out = local.new_empty((n_frames,) + tuple(local.shape[1:]))
    for r in range(world):
        idx = frames_of_rank(n_frames, r, world)
        out[idx] = parts[r, : len(idx)]
    return out


def render_sweep(render_frame: Callable[[int], torch.Tensor], n_frames: int, rank: int = 0, world: int = 1,
                 gather: bool = True) -> torch.Tensor:
    """Render frames [0, n_frames) frame-parallel. `render_frame(f)` returns this rank's (H, W, 9)
    buffer for frame f. Returns all frames in order on every rank when gather=True, else the
    local stack."""
    mine = frames_of_rank(n_frames, rank, world)
    bufs: Sequence[torch.Tensor] = [render_frame(f) for f in mine]
    if bufs:
        local = torch.stack(list(bufs), 0)
    else:
        probe = render_frame(0)
        local = probe.new_zeros((0,) + tuple(probe.shape))
    return gather_frames(local, n_frames, rank, world) if gather else local

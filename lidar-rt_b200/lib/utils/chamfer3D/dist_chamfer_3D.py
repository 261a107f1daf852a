"""Drop-in for the reference's Chamfer module, /root/reference/lib/utils/chamfer3D/dist_chamfer_3D.py.

Same names and call contract — ``chamfer_3DDist()(input1, input2) -> (dist1, dist2, idx1, idx2)`` with
``chamfer_3DFunction`` underneath (:32-76) — used by train.py:197-207, eval.py:355-360 and metric_utils.py:449-457.
The reference JIT-compiles an O(n·m) scan (chamfer3D.cu); this module calls ``lrt_chamfer_forward`` /
``lrt_chamfer_backward`` of the C-ABI library (Morton-sorted point hierarchies, csrc/lrt_chamfer.cu) and returns
bit-identical distances and indices. GPU tensors only, as in the reference; no fallback.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.autograd import Function

from lidar_rt_b200 import native

_ctx: dict = {}


def _context(device: torch.device) -> native.Context:
    """One native context (workspace) per device, created on first use."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    c = _ctx.get(key)
    if c is None:
        c = _ctx[key] = native.Context(torch.device("cuda", key))
    return c


class chamfer_3DFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        batchsize, n, dim = xyz1.size()
        assert dim == 3, "Wrong last dimension for the chamfer distance 's input! Check with .size()"
        _, m, dim = xyz2.size()
        assert dim == 3, "Wrong last dimension for the chamfer distance 's input! Check with .size()"
        nctx = _context(xyz1.device)
        dist1, dist2, idx1, idx2 = nctx.chamfer_forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        nctx = _context(xyz1.device)
        gradxyz1, gradxyz2 = nctx.chamfer_backward(xyz1, xyz2, graddist1.contiguous(), graddist2.contiguous(), idx1, idx2)
        return gradxyz1, gradxyz2


class chamfer_3DDist(nn.Module):
    def __init__(self):
        super(chamfer_3DDist, self).__init__()

    def forward(self, input1, input2):
        input1 = input1.contiguous()
        input2 = input2.contiguous()
        return chamfer_3DFunction.apply(input1, input2)

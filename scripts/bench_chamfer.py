"""Per-kernel times of the Chamfer path on one LiDAR frame's clouds (live CUDA events, LRT_OPT_KERNEL_TIMING)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lidar-rt_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from lidar_rt_b200 import native
from oracle.run_ref_chamfer import lidar_clouds      # case generator only (no oracle call)

n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 169600
pr, gt = lidar_clouds(n_rays, seed=11)
a, c = torch.as_tensor(pr[None]).cuda(), torch.as_tensor(gt[None]).cuda()
ctx = native.Context("cuda:0")
g1 = torch.randn(1, pr.shape[0], device="cuda"); g2 = torch.randn(1, gt.shape[0], device="cuda")
for _ in range(3):
    d1, d2, i1, i2 = ctx.chamfer_forward(a, c); ctx.chamfer_backward(a, c, g1, g2, i1, i2)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
K = 20
fw = bw = 0.0
for _ in range(K):
    e0.record(); d1, d2, i1, i2 = ctx.chamfer_forward(a, c); e1.record(); ctx.chamfer_backward(a, c, g1, g2, i1, i2); e2.record()
    torch.cuda.synchronize(); fw += e0.elapsed_time(e1) / K; bw += e1.elapsed_time(e2) / K
ctx.set_option(native.OPT_KERNEL_TIMING, 1)
ctx.kernel_times()
for _ in range(K):
    d1, d2, i1, i2 = ctx.chamfer_forward(a, c); ctx.chamfer_backward(a, c, g1, g2, i1, i2)
kt = {k: round(v[0] / K, 4) for k, v in ctx.kernel_times().items()}
# the same clouds with the second one shuffled: no index correspondence for the starting bound to use
perm = torch.randperm(gt.shape[0], device="cuda")
cs = c[:, perm].contiguous()
for _ in range(3):
    ctx.chamfer_forward(a, cs)
ctx.kernel_times()
for _ in range(K):
    ctx.chamfer_forward(a, cs)
ks = {k: round(v[0] / K, 4) for k, v in ctx.kernel_times().items()}
print(json.dumps({"n": pr.shape[0], "m": gt.shape[0], "forward_ms": round(fw, 4), "backward_ms": round(bw, 4), "kernels_ms": kt,
                  "shuffled_second_cloud_kernels_ms": ks}))

#!/usr/bin/env python
"""oracle/run_ref_train.py — TEST / BASELINE INFRASTRUCTURE (GPU box only; never imported by the product).

BASELINE.json config #5, the REFERENCE side: one train.py iteration (train.py:125-218) assembled from the reference's own pieces,
on the same synthetic Waymo-dynamic scene scripts/train_loop.py uses, timed on the same B200:

    accessor loop + concatenations        lib/gaussian_renderer/__init__.py:76-134    (this repository's op-for-op restatement `_assemble`:
                                                                                       pinned to the reference's statements by tests/test_prepare_golden.py)
    build2DRectangle                      lib/utils/primitive_utils.py:182-224        (oracle/run_ref_optix.proxy_mesh, same torch ops)
    Tracer.build_acceleration_structure   the UNMODIFIED reference `_C` on OptiX      (oracle/_ref_optix, built by oracle/build_ref_optix.sh)
    _Tracer forward / backward            diff_lidar_tracer/__init__.py:13-136        (same 20 / 22 argument calls)
    Chamfer term                          the UNMODIFIED reference extension           (oracle/_ref_chamfer/chamfer_3D.so)
    optimizer                             one torch.optim.Adam(l, lr=0.0, eps=1e-15) with six groups PER ASSET, stepped one after the
                                          other (gaussian_model.py:186-201, gs_loader.py:243-298)
    python oracle/run_ref_train.py OUT.json [--gaussians 2000000] [--actors 40] [--iters 12]
Densification, logging and data loading are left out on both sides.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "lidar-rt_b200"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--gaussians", type=int, default=2_000_000)
    ap.add_argument("--actors", type=int, default=40)
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=16)
    a = ap.parse_args()
    import torch
    import run_ref_optix as rro
    import run_ref_chamfer as rrc
    from lidar_rt_b200 import synthetic as syn
    from lidar_rt_b200.scene import GaussianAsset, LidarSensor
    import lib.gaussian_renderer as gr

    dev = torch.device("cuda", 0)
    tr = rro.RefTracer()                            # _C.OptiXStateWrapper on libnvoptix
    cham = rrc.load_ref()
    P = a.gaussians + a.actors * 10000
    sc = syn.make_street_scene(P, seed=1, n_actors=a.actors, per_actor=10000)
    assets = []
    for k in [-1] + list(range(a.actors)):
        m = sc.actor_id == k
        sub = syn.Scene(sc.means[m], sc.scales[m], sc.rots[m], sc.opac[m], sc.shs[m], sc.actor_id[m], 3)
        poses = None
        if k >= 0:
            poses = {f: (torch.tensor(syn.actor_transform(k, f)[1], device=dev), torch.tensor([[1.0, 0, 0, 0]], device=dev)) for f in range(a.frames)}
        assets.append(GaussianAsset(sub, device=dev, actor_poses=poses))
    sensor = LidarSensor(device=dev)
    for f in range(a.frames):
        sensor.add_frame(f, syn.sensor_pose(f))
    H, W = sensor.H, sensor.W
    bg = torch.tensor([0.0, 0.0, 1.0], device=dev)
    empty = torch.Tensor([]).cuda()
    eye = torch.eye(4, device=dev)

    class RefTracerFn(torch.autograd.Function):     # diff_lidar_tracer/__init__.py:13-136, call for call
        @staticmethod
        def forward(ctx, ray_o, ray_d, vertices, means3D, shs, opacities, scales, rotations):
            out_f, out_u, accum = tr.C.trace_surfels(tr.ctx, True, ray_o, ray_d, vertices, bg, means3D, shs, 3, empty, opacities, scales, 1.0,
                                                     rotations, empty, eye, eye, torch.zeros(3, device=dev), False, False)
            ctx.save_for_backward(ray_o, ray_d, vertices, means3D, shs, opacities, scales, rotations, out_f, out_u)
            return out_f, accum

        @staticmethod
        def backward(ctx, g, _):
            ray_o, ray_d, vertices, means3D, shs, opacities, scales, rotations, out_f, out_u = ctx.saved_tensors
            gm, gsh, _, gop, gsc, grot, _, _ = tr.C.trace_surfels_backward(tr.ctx, ray_o, ray_d, vertices, bg, means3D, shs, 3, empty, opacities, scales,
                                                                           1.0, rotations, empty, eye, eye, torch.zeros(3, device=dev), False, False,
                                                                           out_f, out_u, g.contiguous())
            return None, None, None, gm, gsh, gop, gsc, grot

    class RefChamfer(torch.autograd.Function):      # lib/utils/chamfer3D/dist_chamfer_3D.py:32-76
        @staticmethod
        def forward(ctx, xyz1, xyz2):
            b, n, _ = xyz1.size(); m = xyz2.size(1)
            d1 = torch.zeros(b, n, device=dev); d2 = torch.zeros(b, m, device=dev)
            i1 = torch.zeros(b, n, device=dev, dtype=torch.int32); i2 = torch.zeros(b, m, device=dev, dtype=torch.int32)
            cham.forward(xyz1, xyz2, d1, d2, i1, i2)
            ctx.save_for_backward(xyz1, xyz2, i1, i2)
            return d1, d2, i1, i2

        @staticmethod
        def backward(ctx, g1, g2, _a, _b):
            xyz1, xyz2, i1, i2 = ctx.saved_tensors
            gx1 = torch.zeros_like(xyz1); gx2 = torch.zeros_like(xyz2)
            cham.backward(xyz1, xyz2, gx1, gx2, g1.contiguous(), g2.contiguous(), i1, i2)
            return gx1, gx2

    def render(f):
        rays_o, rays_d = sensor.get_range_rays(f)
        means3D, opacity, scales, rotations, shs = gr._assemble(f, assets, True, False)                  # :76-134
        verts, faces = rro.proxy_mesh(means3D, scales, rotations, opacity)                                # :142
        tr.C.build_acceleration_structure(tr.ctx, verts, faces, True)                                     # :145
        out, accum = RefTracerFn.apply(rays_o.contiguous(), rays_d, verts, means3D, shs, opacity, scales, rotations)
        prob = torch.softmax(torch.cat([out[..., 1:2], out[..., 2:3]], dim=-1), dim=-1)                   # :163-173
        return {"depth": out[..., 3:4], "intensity": out[..., 0:1], "raydrop": prob[..., 1:2]}

    gt = {}
    with torch.no_grad():
        for f in range(a.frames):
            pkg = render(f)
            gt[f] = (pkg["depth"] * (1 + 0.02 * torch.randn_like(pkg["depth"])), (pkg["intensity"] + 0.05 * torch.randn_like(pkg["intensity"])).clamp(0, 1),
                     (pkg["raydrop"] > 0.5).float())
    lrs = dict(_xyz=1.6e-4, _features_dc=2.5e-3, _features_rest=1.25e-4, _opacity=0.05, _scaling=5e-3, _rotation=1e-3)     # configs/exp.yaml
    opts = [torch.optim.Adam([{"params": [getattr(x, n)], "lr": lr, "name": n} for n, lr in lrs.items()], lr=0.0, eps=1e-15) for x in assets]
    rng = np.random.default_rng(0)

    def iteration():
        f = int(rng.integers(0, a.frames))
        pkg = render(f)
        d_gt, i_gt, r_gt = gt[f]
        loss = (pkg["depth"] - d_gt).abs().mean() * 0.1 + (pkg["intensity"] - i_gt).abs().mean() + \
            torch.nn.functional.binary_cross_entropy(pkg["raydrop"].clamp(1e-6, 1 - 1e-6), r_gt) * 0.1
        mask = r_gt[..., 0] < 0.5
        gt_pts = sensor.inverse_projection_with_range(f, d_gt, mask)
        pred_pts = sensor.inverse_projection_with_range(f, pkg["depth"], mask)
        d1, d2, _, _ = RefChamfer.apply(pred_pts[None, ...], gt_pts[None, ...])
        loss = loss + 0.01 * (d1 + d2).mean() * 0.5
        loss.backward()
        for o in opts:                               # gs_loader.py:243-298: every asset's optimizer, one after the other
            o.step()
            o.zero_grad(set_to_none=True)
        return loss

    for _ in range(a.warmup):
        iteration()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(a.iters):
        loss = iteration()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    rep = {"what": "reference-equivalent train.py iteration: accessor loop + build2DRectangle + OptiX accel rebuild + diff-lidar-tracer fwd/bwd (unmodified _C on "
                   "OptiX) + reference chamfer3D + per-asset torch.optim.Adam", "P": P, "actors": a.actors, "rays": H * W, "iters": a.iters,
           "it_per_s": a.iters / dt, "ms_per_it": 1e3 * dt / a.iters, "final_loss": float(loss), "gpu": torch.cuda.get_device_name(0)}
    json.dump(rep, open(a.out, "w"), indent=1)
    print(json.dumps(rep))


if __name__ == "__main__":
    sys.exit(main())

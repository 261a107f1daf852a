#!/usr/bin/env python
"""oracle/run_ref_chamfer.py — TEST INFRASTRUCTURE (GPU box only; never imported by the product).

Runs the UNMODIFIED reference Chamfer extension (oracle/_ref_chamfer/chamfer_3D.so, built by
oracle/build_ref_chamfer.sh from /root/reference/lib/utils/chamfer3D) through the two native calls its Python
wrapper makes (dist_chamfer_3D.py:55,71):

    chamfer_3D.forward(xyz1, xyz2, dist1, dist2, idx1, idx2); chamfer_3D.backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2)

    python oracle/run_ref_chamfer.py golden OUT.npz   # the reference's outputs for the seeded cases of chamfer_cases()
    python oracle/run_ref_chamfer.py bench  OUT.json  # LiDAR-frame-sized clouds: reference ms, this repo's ms, and
                                                      # bit-exact comparison of every distance and index
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "lidar-rt_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def lidar_clouds(n_rays: int, seed: int = 0, noise: float = 0.05, drop: float = 0.15):
    """Two LiDAR-frame-shaped clouds, the way train.py:198-203 makes them: the same rays back-projected with the
    ground-truth range and with a predicted range (ground truth + noise), for the rays the mask keeps."""
    from lidar_rt_b200 import synthetic as syn
    W = max(1, n_rays // syn.WAYMO_H)
    o, d = syn.lidar_rays(syn.WAYMO_H, W, syn.waymo_inclinations(), syn.sensor_pose(3))
    d = d.reshape(-1, 3).astype(np.float32)
    rng = np.random.default_rng(seed)
    # ranges: ground plane 2 m below the sensor for downward beams, walls at 8-40 m otherwise
    down = d[:, 2] < -0.02
    r = np.where(down, np.minimum(2.0 / np.maximum(-d[:, 2], 1e-3), 75.0), rng.uniform(8.0, 40.0, d.shape[0])).astype(np.float32)
    keep = rng.random(d.shape[0]) > drop
    gt = (o.reshape(-1, 3)[:1] + d * r[:, None])[keep].astype(np.float32)
    pr = (o.reshape(-1, 3)[:1] + d * (r + rng.normal(0, noise, r.shape).astype(np.float32))[:, None])[keep].astype(np.float32)
    return pr, gt


def chamfer_cases():
    """name -> (xyz1 (b,n,3), xyz2 (b,m,3)); deterministic."""
    rng = np.random.default_rng(1234)
    f = lambda *s: rng.normal(size=s).astype(np.float32)
    cases = {}
    cases["random_3000x2500"] = (f(1, 3000, 3), f(1, 2500, 3))
    cases["batch2_700x1300"] = (f(2, 700, 3) * 3, f(2, 1300, 3) * 3 + 0.5)
    # exact ties: small integer lattices with many duplicates (the lowest index must win)
    cases["lattice_ties"] = (rng.integers(0, 6, (1, 2000, 3)).astype(np.float32), rng.integers(0, 6, (1, 1500, 3)).astype(np.float32))
    cases["one_one"] = (f(1, 1, 3), f(1, 1, 3))
    cases["five_three"] = (f(1, 5, 3), f(1, 3, 3))
    cases["nine_eight"] = (f(1, 9, 3), f(1, 8, 3))
    # sizes around the reference's 512-point batches, with the same point duplicated on both sides of a batch boundary
    for m in (511, 512, 513, 1024, 1025):
        c = f(1, m, 3)
        if m > 600:
            c[0, 600] = c[0, 10]
        c[0, m - 1] = c[0, 3]
        cases[f"batch_edge_{m}"] = (np.concatenate([c[:, ::7] + 1e-3, c[:, [3, 10]]], 1), c)
    pr, gt = lidar_clouds(8192, seed=5)
    cases["lidar_8192"] = (pr[None], gt[None])
    # identical clouds (distance 0 everywhere, index = self unless an earlier duplicate exists) and a far outlier
    a = f(1, 1000, 3); a[0, 500] = a[0, 20]; a[0, 999] = 1e4
    cases["self"] = (a, a.copy())
    return cases


def load_ref():
    import torch  # noqa: F401
    path = os.path.join(HERE, "_ref_chamfer", "chamfer_3D.so")
    spec = importlib.util.spec_from_file_location("chamfer_3D", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class RefChamfer:
    """chamfer_3DFunction.forward / .backward of the reference (dist_chamfer_3D.py:34-76), call for call."""

    def __init__(self):
        self.C = load_ref()

    def forward(self, xyz1, xyz2):
        import torch
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        dev = xyz1.device
        dist1 = torch.zeros(b, n).to(dev); dist2 = torch.zeros(b, m).to(dev)
        idx1 = torch.zeros(b, n).type(torch.IntTensor).to(dev); idx2 = torch.zeros(b, m).type(torch.IntTensor).to(dev)
        self.C.forward(xyz1, xyz2, dist1, dist2, idx1, idx2)
        return dist1, dist2, idx1, idx2

    def backward(self, xyz1, xyz2, g1, g2, idx1, idx2):
        import torch
        ga = torch.zeros(xyz1.size()).to(xyz1.device); gc = torch.zeros(xyz2.size()).to(xyz1.device)
        self.C.backward(xyz1, xyz2, ga, gc, g1.contiguous(), g2.contiguous(), idx1, idx2)
        return ga, gc


def case_grads(name, b, n, m):
    rng = np.random.default_rng(sum(map(ord, name)))
    return rng.normal(size=(b, n)).astype(np.float32), rng.normal(size=(b, m)).astype(np.float32)


def golden(out_path):
    import torch
    ref = RefChamfer()
    res = {}
    for name, (a, c) in chamfer_cases().items():
        ta, tc = torch.as_tensor(a).cuda(), torch.as_tensor(c).cuda()
        d1, d2, i1, i2 = ref.forward(ta, tc)
        g1, g2 = case_grads(name, a.shape[0], a.shape[1], c.shape[1])
        ga, gc = ref.backward(ta, tc, torch.as_tensor(g1).cuda(), torch.as_tensor(g2).cuda(), i1, i2)
        torch.cuda.synchronize()
        for k, v in (("a", a), ("c", c), ("d1", d1), ("d2", d2), ("i1", i1), ("i2", i2), ("g1", g1), ("g2", g2), ("ga", ga), ("gc", gc)):
            res[f"{name}/{k}"] = v.cpu().numpy() if hasattr(v, "cpu") else v
        print(f"{name}: n={a.shape[1]} m={c.shape[1]} mean d1 {float(d1.mean()):.4g}")
    np.savez_compressed(out_path, **res)
    print("wrote", out_path)


def bench(out_path, n_rays=169600, iters=5):
    import torch
    from lib.utils.chamfer3D.dist_chamfer_3D import chamfer_3DFunction
    from lidar_rt_b200 import native
    ref = RefChamfer()
    pr, gt = lidar_clouds(n_rays, seed=11)
    ta, tc = torch.as_tensor(pr[None]).cuda(), torch.as_tensor(gt[None]).cuda()
    rec = {"n": int(pr.shape[0]), "m": int(gt.shape[0]), "gpu": torch.cuda.get_device_name(0)}

    def timed(fn):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(iters):
            e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts)), out

    g1 = torch.randn(1, pr.shape[0], device="cuda"); g2 = torch.randn(1, gt.shape[0], device="cuda")
    rec["reference_forward_ms"], (rd1, rd2, ri1, ri2) = timed(lambda: ref.forward(ta, tc))
    rec["reference_backward_ms"], (rga, rgc) = timed(lambda: ref.backward(ta, tc, g1, g2, ri1, ri2))
    nctx = native.Context("cuda:0")
    rec["ours_forward_ms"], (d1, d2, i1, i2) = timed(lambda: nctx.chamfer_forward(ta, tc))
    rec["ours_backward_ms"], (ga, gc) = timed(lambda: nctx.chamfer_backward(ta, tc, g1, g2, i1, i2))
    rec["dist_bit_exact"] = bool(torch.equal(d1, rd1) and torch.equal(d2, rd2))
    rec["idx_exact"] = bool(torch.equal(i1, ri1) and torch.equal(i2, ri2))
    rec["dist_mismatches"] = int((d1 != rd1).sum() + (d2 != rd2).sum())
    rec["idx_mismatches"] = int((i1 != ri1).sum() + (i2 != ri2).sum())
    rec["grad_max_rel_err"] = float(max((ga - rga).abs().max() / rga.abs().max(), (gc - rgc).abs().max() / rgc.abs().max()))
    # through the drop-in autograd surface as train.py:205-206 uses it
    pa = ta.clone().requires_grad_(True)
    da, db, _, _ = chamfer_3DFunction.apply(pa, tc)
    ((da + db).mean() * 0.5).backward()
    rec["autograd_grad_finite"] = bool(torch.isfinite(pa.grad).all())
    rec["speedup_forward"] = rec["reference_forward_ms"] / rec["ours_forward_ms"]
    rec["speedup_fwd_bwd"] = (rec["reference_forward_ms"] + rec["reference_backward_ms"]) / (rec["ours_forward_ms"] + rec["ours_backward_ms"])
    print(json.dumps(rec, indent=1))
    with open(out_path, "w") as f:
        json.dump(rec, f, indent=1)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "golden"
    if mode == "golden":
        golden(sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "chamfer_ref_b200.npz"))
    elif mode == "bench":
        bench(sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "chamfer_bench.json"))
    else:
        raise SystemExit(__doc__)

#!/usr/bin/env python
"""oracle/run_ref_optix.py — TEST INFRASTRUCTURE (GPU box only; never imported by the product).

Runs the UNMODIFIED reference tracer (built by oracle/build_ref_optix.sh into oracle/_ref_optix/:
its pybind11 module `_C` + forward.ptx / backward.ptx) on the real OptiX runtime in the driver
(libnvoptix.so.1), through the same native calls the reference's Python wrapper makes
(submodules/diff-lidar-tracer/diff_lidar_tracer/__init__.py:15-136, 164-171):

    _C.OptiXStateWrapper(pkg_dir); _C.build_acceleration_structure(ctx, vertices, triangles, rebuild)
    _C.trace_surfels(...20 args...) ; _C.trace_surfels_backward(...22 args...)

with the proxy mesh of lib/utils/primitive_utils.py:182-224 (`build2DRectangle`, restated in
`proxy_mesh()` below with the same torch operations).

    python oracle/run_ref_optix.py probe            # is libnvoptix there, does the context come up
    python oracle/run_ref_optix.py golden OUT.npz   # real-OptiX outputs for the inputs of tests/golden/*
    python oracle/run_ref_optix.py bench  OUT.json  # B1: Mrays/s of the reference on the BASELINE workload,
                                                    # and full-size parity of this repo's library against it
    python oracle/run_ref_optix.py fullsize OUT.npz [P frame seed]   # full-size goldens: outputs on every 16th ray +
                                                    # sparse gradients for dL_dout on every 128th ray

Each mode runs in a child process (a missing OptiX makes the reference segfault: its failed
optixInit() is only printed, optix_wrapper.cpp:186).
"""
from __future__ import annotations

import ctypes
import glob
import importlib.util
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(HERE, "_ref_optix")
for p in (ROOT, os.path.join(ROOT, "lidar-rt_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

BG = np.array([0.0, 0.0, 1.0], np.float32)


def find_nvoptix():
    cands = []
    for d in ("/usr/local/nvidia/lib", "/usr/local/nvidia/lib64", "/usr/lib/x86_64-linux-gnu", "/usr/lib64", "/usr/lib"):
        cands += sorted(glob.glob(os.path.join(d, "libnvoptix.so*")))
    return cands


def load_ref():
    """dlopen libnvoptix by path first (optixInit() dlopens it by soname: glibc then reuses the loaded
    object even when the directory is not on the loader path), then import the reference's module."""
    import torch  # noqa: F401  (the module links against libtorch)
    libs = find_nvoptix()
    if not libs:
        raise RuntimeError("libnvoptix.so.1 not found on this machine")
    ctypes.CDLL(libs[0], mode=ctypes.RTLD_GLOBAL)
    spec = importlib.util.spec_from_file_location("_C", os.path.join(PKG, "_C.so"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m, libs[0]


def proxy_mesh(means, scales, rots, opac):
    """build2DRectangle (primitive_utils.py:182-224) + build_rotation (general_utils.py:176-197)."""
    import torch
    P = means.shape[0]
    dev = means.device
    local = torch.tensor([[-1, 1, 0], [-1, -1, 0], [1, 1, 0], [1, -1, 0]], device=dev).repeat(P, 1, 1).float()
    S = torch.zeros((P, 3, 3), dtype=torch.float, device=dev)
    factor = (torch.sqrt(2 * torch.log(opac / (1.0 / 255.0))) + 0.01).reshape(P)
    S[:, 0, 0] = scales[:, 0] * factor
    S[:, 1, 1] = scales[:, 1] * factor
    S[:, 2, 2] = 1
    r = rots
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    Rm = torch.zeros((P, 3, 3), device=dev)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    Rm[:, 0, 0] = 1 - 2 * (y * y + z * z); Rm[:, 0, 1] = 2 * (x * y - w * z); Rm[:, 0, 2] = 2 * (x * z + w * y)
    Rm[:, 1, 0] = 2 * (x * y + w * z); Rm[:, 1, 1] = 1 - 2 * (x * x + z * z); Rm[:, 1, 2] = 2 * (y * z - w * x)
    Rm[:, 2, 0] = 2 * (x * z - w * y); Rm[:, 2, 1] = 2 * (y * z + w * x); Rm[:, 2, 2] = 1 - 2 * (x * x + y * y)
    verts = local @ (Rm @ S).permute(0, 2, 1) + means.unsqueeze(1).repeat(1, 4, 1)
    base = torch.tensor([[0, 1, 2], [2, 3, 1]], device=dev)
    faces = (base + torch.arange(0, P * 4, step=4, device=dev).view(P, 1, 1)).int()
    return verts.view(-1, 3).contiguous(), faces.view(-1, 3).contiguous()


class RefTracer:
    """The reference's Tracer, call for call."""

    def __init__(self):
        import torch
        self.torch = torch
        self.C, self.lib = load_ref()
        self.ctx = self.C.OptiXStateWrapper(PKG)
        self.empty = torch.Tensor([]).cuda()
        self.eye = torch.eye(4, device="cuda")

    def build(self, means, scales, rots, opac):
        # nan_to_num is NOT applied: opacity < 1/255 gives NaN vertices exactly as in the reference
        self.vertices, self.triangles = proxy_mesh(means, scales, rots, opac)
        self.C.build_acceleration_structure(self.ctx, self.vertices, self.triangles, True)

    def forward(self, ray_o, ray_d, bg, means, scales, rots, opac, shs, D):
        t = self.torch
        e = self.empty
        return self.C.trace_surfels(self.ctx, True, ray_o, ray_d, self.vertices, bg, means, shs, int(D), e, opac, scales, 1.0,
                                    rots, e, self.eye, self.eye, t.zeros(3, device="cuda"), False, False)

    def backward(self, ray_o, ray_d, bg, means, scales, rots, opac, shs, D, out_f, out_u, dL):
        t = self.torch
        e = self.empty
        return self.C.trace_surfels_backward(self.ctx, ray_o, ray_d, self.vertices, bg, means, shs, int(D), e, opac, scales, 1.0,
                                             rots, e, self.eye, self.eye, t.zeros(3, device="cuda"), False, False, out_f, out_u, dL)


def run_case(tr: RefTracer, o, d, sc, D, dL=None):
    """o (1,3) or (H,W,3); d (H,W,3) or (R,3). Returns numpy dict like oracle/make_golden.run_ref."""
    import torch
    cu = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
    d = np.asarray(d, np.float32)
    if d.ndim == 2:
        d = d.reshape(1, -1, 3)
    Hh, Ww = d.shape[:2]
    o = np.asarray(o, np.float32)
    ro = cu(np.broadcast_to(o.reshape(-1, 3)[0] if o.size == 3 else o.reshape(Hh, Ww, 3), (Hh, Ww, 3)))
    rd = cu(d)
    means, scales, rots, opac, shs = (cu(sc[k]) for k in ("means", "scales", "rots", "opac", "shs"))
    bg = cu(BG)
    tr.build(means, scales, rots, opac)
    out_f, out_u, accum = tr.forward(ro, rd, bg, means, scales, rots, opac, shs, D)
    res = dict(out=out_f.reshape(-1, 9).cpu().numpy(), accum_w=accum.cpu().numpy())
    if dL is not None:
        g = tr.backward(ro, rd, bg, means, scales, rots, opac, shs, D, out_f, out_u, cu(np.asarray(dL, np.float32).reshape(Hh, Ww, 9)))
        gm, gsh, _, gop, gsc, grot, _, _ = g
        res.update(g_means=gm.cpu().numpy(), g_shs=gsh.cpu().numpy(), g_opac=gop.cpu().numpy().reshape(-1),
                   g_scales=gsc.cpu().numpy(), g_rots=grot.cpu().numpy())
    torch.cuda.synchronize()
    return res


# ------------------------------------------------------------------------------------------ modes
def mode_probe():
    tr = RefTracer()
    print(json.dumps({"nvoptix": tr.lib, "context": "ok"}))


def mode_golden(out_path):
    """The inputs of tests/golden/ref_kat.npz, ref_scene_small.npz and config #1, through real OptiX."""
    from lidar_rt_b200 import synthetic as syn
    tr = RefTracer()
    G = os.path.join(ROOT, "tests", "golden")
    out = {}
    kat = np.load(os.path.join(G, "ref_kat.npz"))
    rep = {}
    for name in [str(n) for n in kat["names"]]:
        sc = {k: kat[f"{name}/{k}"] for k in ("means", "scales", "rots", "opac", "shs")}
        res = run_case(tr, kat[f"{name}/ray_o"], kat[f"{name}/ray_d"], sc, int(kat[f"{name}/D"]), kat[f"{name}/dL"])
        for k, v in res.items():
            out[f"kat/{name}/{k}"] = v
        rep[name] = float(np.abs(res["out"] - kat[f"{name}/out"]).max())
    sm = np.load(os.path.join(G, "ref_scene_small.npz"))
    sc = {k: sm[k] for k in ("means", "scales", "rots", "opac", "shs")}
    res = run_case(tr, sm["ray_o"], sm["ray_d"], sc, int(sm["D"]), sm["dL"])
    for k, v in res.items():
        out[f"small/{k}"] = v
    rep["scene_small"] = float(np.abs(res["out"] - sm["out"]).max())
    rep["scene_small_grad_rel"] = {k: float(np.abs(res[k] - sm[k].reshape(res[k].shape)).max() / (np.abs(sm[k]).max() + 1e-30))
                                   for k in ("g_means", "g_shs", "g_opac", "g_scales", "g_rots")}
    scn = syn.make_street_scene(10000, seed=0)
    o, d = syn.ray_patch(64, 64)
    sc = dict(means=scn.means, scales=scn.scales, rots=scn.rots, opac=scn.opac, shs=scn.shs)
    res = run_case(tr, o, d, sc, 3)
    c1 = np.load(os.path.join(G, "ref_cfg1_forward.npz"))
    out["cfg1/out"] = res["out"]; out["cfg1/accum_w"] = res["accum_w"]
    rep["cfg1"] = float(np.abs(res["out"] - c1["out"]).max())
    np.savez_compressed(out_path, **out)
    print(json.dumps({"golden": out_path, "max_abs_diff_vs_host_compiled_reference": rep}))


def mode_bench(out_path, P=2_000_000, steps=10, warmup=3, seed=1):
    """B1: the reference on the BASELINE workload (same inputs as bench.py), CUDA-event timed; then this
    repo's library on the same frame, compared output for output."""
    import torch
    from lidar_rt_b200 import native, synthetic as syn
    H, W = 64, 2650
    R = H * W
    tr = RefTracer()
    sc = syn.make_street_scene(P, seed=seed)
    inc = syn.waymo_inclinations()
    cu = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
    means, scales, rots, opac, shs = map(cu, (sc.means, sc.scales, sc.rots, sc.opac, sc.shs))
    bg = cu(BG)
    rng = np.random.default_rng(1000)
    frames = []
    for f in range(steps + warmup):
        o, d = syn.lidar_rays(H, W, inc, syn.sensor_pose(f))
        dL = np.zeros((H, W, 9), np.float32); dL[..., :4] = rng.standard_normal((H, W, 4)).astype(np.float32)
        frames.append((cu(np.broadcast_to(o.reshape(3), (H, W, 3))), cu(d), cu(dL), cu(o)))
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(steps + warmup)]
    for i, (ro, rd, dL, _) in enumerate(frames):
        ev[i][0].record()
        v, t = proxy_mesh(means, scales, rots, opac)
        ev[i][1].record()
        tr.vertices, tr.triangles = v, t
        tr.C.build_acceleration_structure(tr.ctx, v, t, True)          # rebuild=True every call: gaussian_renderer/__init__.py:145
        ev[i][2].record()
        out_f, out_u, accum = tr.forward(ro, rd, bg, means, scales, rots, opac, shs, 3)
        ev[i][3].record()
        g = tr.backward(ro, rd, bg, means, scales, rots, opac, shs, 3, out_f, out_u, dL)
        ev[i][4].record()
    torch.cuda.synchronize()
    span = lambda a, b: float(np.median([e[a].elapsed_time(e[b]) for e in ev[warmup:]]))
    rep = {"P": P, "rays": R, "steps": steps, "warmup": warmup,
           "ms": {"proxy_mesh": span(0, 1), "accel_build": span(1, 2), "forward": span(2, 3), "backward": span(3, 4), "step": span(0, 4)}}
    rep["mrays_per_s_fwd_bwd"] = R / (rep["ms"]["step"] * 1e-3) / 1e6
    print(json.dumps({"b1": rep}), flush=True)

    # ---- full-size parity of this repo's library against the real reference, last frame
    ro, rd, dL, oc = frames[-1]
    ref_out = out_f.reshape(-1, 9); ref_acc = accum
    ref_g = dict(means=g[0], shs=g[1], opac=g[3].reshape(-1), scales=g[4], rots=g[5])
    ctx = native.Context("cuda:0")
    ctx.build(means, scales, rots, opac)
    f = ctx.forward(oc, rd, bg, means, scales, rots, opac, shs, 3)
    gg = ctx.backward(oc, rd, bg, means, scales, rots, opac, shs, 3, f["out"], dL, hits=f)
    torch.cuda.synchronize()
    mine = f["out"].reshape(-1, 9)
    par = {}
    names = ["intensity", "hit_logit", "drop_logit", "depth", "accum", "n0", "n1", "n2", "final_T"]
    diff = (mine - ref_out).abs()
    for c, n in enumerate(names):
        if n.startswith("n"):
            continue
        dc = diff[:, c]
        par[n] = {"max_abs": float(dc.max()), "p999_abs": float(torch.quantile(dc, 0.999)), "mean_abs": float(dc.mean()),
                  "rays_over_1e-4": int((dc > 1e-4).sum())}
    rel_depth = diff[:, 3] / (1.0 + ref_out[:, 3].abs())
    par["depth_rel"] = {"max": float(rel_depth.max()), "rays_over_1e-4": int((rel_depth > 1e-4).sum())}
    par["accum_w"] = {"max_abs": float((f["accum_w"] - ref_acc).abs().max()), "ref_max": float(ref_acc.abs().max())}
    for k in ("means", "shs", "opac", "scales", "rots"):
        a = gg[k].reshape(ref_g[k].shape); b = ref_g[k]
        par["grad_" + k] = {"max_abs": float((a - b).abs().max()), "ref_max": float(b.abs().max()),
                            "rel_l2": float((a - b).norm() / (b.norm() + 1e-30))}
    rep["parity_full_size"] = par
    rep["parity_frame"] = steps + warmup - 1
    if os.environ.get("LRT_REF_OPTIX_DUMP"):          # both renders of the compared frame, for offline diagnosis
        np.savez_compressed(os.environ["LRT_REF_OPTIX_DUMP"], ref_out=ref_out.cpu().numpy(), out=mine.cpu().numpy(),
                            hit_cnt=f["hit_cnt"].cpu().numpy(), frame=np.int32(steps + warmup - 1))
    ctx.close()
    json.dump(rep, open(out_path, "w"), indent=1)
    print(json.dumps({"parity_full_size": par}))


def fullsize_subsets(R, out_stride=16, grad_stride=128):
    """The fixed ray subsets of the full-size goldens: outputs on every `out_stride`-th ray, upstream gradients on every
    `grad_stride`-th ray (zero elsewhere). Shared with tests/test_optix_golden.py through the npz itself."""
    return np.arange(0, R, out_stride, dtype=np.int32), np.arange(0, R, grad_stride, dtype=np.int32)


def fullsize_dL(R, grad_rays, dl_seed):
    """dL_dout of the full-size goldens: N(0,1) on channels 0-3 of the `grad_rays`, zero elsewhere."""
    dL = np.zeros((R, 9), np.float32)
    dL[grad_rays, :4] = np.random.default_rng(dl_seed).standard_normal((len(grad_rays), 4)).astype(np.float32)
    return dL


def mode_fullsize(out_path, P=2_000_000, frame=5, seed=1, out_stride=16, grad_stride=128, dl_seed=4242, sh_keep=8):
    """Full-size golden vectors from the reference on OptiX: one 64 x 2650 frame over P Gaussians, forward outputs on
    every 16th ray, and the GRADIENTS the reference's backward produces when dL_dout is non-zero on every 128th ray only.
    The gradients are stored sparsely (rows of the Gaussians they touch; SH rows for every `sh_keep`-th of those)."""
    from lidar_rt_b200 import synthetic as syn
    H, W = 64, 2650
    R = H * W
    tr = RefTracer()
    scn = syn.make_street_scene(P, seed=seed)
    o, d = syn.lidar_rays(H, W, syn.waymo_inclinations(), syn.sensor_pose(frame))
    out_rays, grad_rays = fullsize_subsets(R, out_stride, grad_stride)
    dL = fullsize_dL(R, grad_rays, dl_seed)
    sc = dict(means=scn.means, scales=scn.scales, rots=scn.rots, opac=scn.opac, shs=scn.shs)
    res = run_case(tr, o, d, sc, 3, dL)
    chans = np.array([0, 1, 2, 3, 4, 8], np.int32)
    touched = np.flatnonzero((np.abs(res["g_means"]).sum(1) + np.abs(res["g_opac"]) + np.abs(res["g_scales"]).sum(1) +
                              np.abs(res["g_rots"]).sum(1) + np.abs(res["g_shs"]).reshape(P, -1).sum(1)) > 0).astype(np.int32)
    sh_rows = touched[::sh_keep]
    out = dict(P=np.int32(P), seed=np.int32(seed), frame=np.int32(frame), ray_index=out_rays, channels=chans,
               out=res["out"][out_rays][:, chans], grad_rays=grad_rays, dl_seed=np.int32(dl_seed),
               out_grad_rays=res["out"][grad_rays], g_index=touched, g_means=res["g_means"][touched], g_opac=res["g_opac"][touched],
               g_scales=res["g_scales"][touched], g_rots=res["g_rots"][touched], g_sh_index=sh_rows, g_shs=res["g_shs"][sh_rows],
               g_norms=np.array([np.linalg.norm(res[k].astype(np.float64)) for k in ("g_means", "g_shs", "g_opac", "g_scales", "g_rots")]),
               accum_w_touched=res["accum_w"][touched])
    np.savez_compressed(out_path, **out)
    print(json.dumps({"fullsize": out_path, "P": P, "touched_gaussians": int(len(touched)), "sh_rows": int(len(sh_rows)),
                      "bytes": os.path.getsize(out_path)}))


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "probe"
    if os.environ.get("_LRT_REF_OPTIX_CHILD") == "1":
        if mode == "probe":
            mode_probe()
        elif mode == "golden":
            mode_golden(sys.argv[2])
        elif mode == "bench":
            mode_bench(sys.argv[2], *(int(a) for a in sys.argv[3:]))
        elif mode == "fullsize":
            mode_fullsize(sys.argv[2], *(int(a) for a in sys.argv[3:]))
        return 0
    if not os.path.exists(os.path.join(PKG, "_C.so")):
        print(json.dumps({"unavailable": "oracle/_ref_optix not built (oracle/build_ref_optix.sh needs /root/reference)"}))
        return 0
    env = dict(os.environ, _LRT_REF_OPTIX_CHILD="1")
    t0 = time.time()
    try:
        rc = subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env,
                            timeout=float(os.environ.get("LRT_REF_OPTIX_TIMEOUT", "600"))).returncode
    except subprocess.TimeoutExpired:
        rc = -999
    print(json.dumps({"mode": mode, "child_rc": rc, "seconds": round(time.time() - t0, 1)}))
    return 0


if __name__ == "__main__":
    sys.exit(main())

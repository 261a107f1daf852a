#!/usr/bin/env python
"""oracle/make_golden.py — TEST INFRASTRUCTURE. Generates tests/golden/*.npz.

Run in the build container (needs /root/reference):  python oracle/make_golden.py

Two sources, both the reference's OWN code:
 1. oracle/_ref — forward.cu / backward.cu compiled unmodified as host C++ over the OptiX stand-in
    (oracle/build_ref.sh). Gives forward outputs, accum weights and gradients for seeded scenes
    and hand-built known-answer cases.
 2. The reference's Python helpers, executed unmodified on the CPU: the function bodies are
    lifted out of their files with `ast` (their modules import open3d / tensorflow / icosphere,
    which do not exist here) and run with `device="cuda"` / `.cuda()` redirected to the CPU:
      lib/utils/sh_utils.py::eval_sh               (SH basis cross-check)
      lib/utils/general_utils.py::build_rotation, quaternion_raw_multiply
      lib/utils/primitive_utils.py::build2DRectangle
      lib/scene/lidar_sensor.py::LiDARSensor.get_range_rays
      lib/scene/gaussian_model.py::GaussianModel.get_world_xyz / get_rotation / get_scaling / get_opacity / get_features
          and the accessor loop + concatenations + rotation composition of lib/gaussian_renderer/__init__.py:68-134
          (statements cut out of raytracing() and executed as they are) -> ref_prepare.npz, values AND leaf gradients
      lib/utils/graphics_utils.py::get_rays                    (pinhole rays of the Camera branch of raytracing())
      lib/scene/gaussian_model.py::GaussianModel.prune_points / densify_and_clone / densify_and_split / densify_and_prune (+ the
          optimizer-state helpers they call), over a real torch.optim.Adam -> ref_densify.npz
The fixtures are small; the GPU box has no /root/reference, so tests read only these files.
"""
from __future__ import annotations

import ast
import contextlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("LIDAR_RT_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lidar-rt_b200"))

from oracle.oracle import Oracle, Ref  # noqa: E402
from lidar_rt_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


# ------------------------------------------------------------------ reference python, on the CPU
@contextlib.contextmanager
def cuda_to_cpu():
    """Redirect device='cuda' factory calls and .cuda() to the CPU while reference code runs."""
    names = ["tensor", "zeros", "ones", "arange", "zeros_like", "ones_like", "empty", "full", "eye"]
    saved = {n: getattr(torch, n) for n in names}
    saved_cuda = torch.Tensor.cuda

    def wrap(fn):
        def inner(*a, **k):
            if str(k.get("device", "")).startswith("cuda"):
                k["device"] = "cpu"
            return fn(*a, **k)
        return inner

    for n in names:
        setattr(torch, n, wrap(saved[n]))
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        for n in names:
            setattr(torch, n, saved[n])
        torch.Tensor.cuda = saved_cuda


def lift(path: str, names: list[str], ns: dict | None = None, strip_decorators: bool = False) -> dict:
    """exec only the named top-level defs / assignments (or methods `Class.method`) of a file."""
    src = open(path).read()
    tree = ast.parse(src)
    picked = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            picked.append(node)
        elif isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id in names for t in node.targets):
            picked.append(node)
        elif isinstance(node, ast.ClassDef):
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and f"{node.name}.{sub.name}" in names:
                    if strip_decorators:
                        sub.decorator_list = []           # @property accessors become plain functions of `self`
                    picked.append(sub)
    ns = ns if ns is not None else {}
    ns.setdefault("torch", torch); ns.setdefault("np", np); ns.setdefault("F", torch.nn.functional)
    exec(compile(ast.Module(picked, []), path, "exec"), ns)
    return ns


def gen_python_fixtures():
    rng = np.random.default_rng(7)
    out = {}
    # eval_sh: sh [..., C, (deg+1)^2], dirs [..., 3]
    ns = lift(f"{REF}/lib/utils/sh_utils.py", ["C0", "C1", "C2", "C3", "C4", "eval_sh"])
    N = 64
    sh = rng.standard_normal((N, 16, 3)).astype(np.float32)          # tracer layout (P, M, 3)
    dirs = rng.standard_normal((N, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    out["sh_coeffs"] = sh; out["sh_dirs"] = dirs
    for deg in range(4):
        n = (deg + 1) ** 2
        res = ns["eval_sh"](deg, torch.from_numpy(sh[:, :n, :]).transpose(1, 2), torch.from_numpy(dirs))
        out[f"sh_eval_deg{deg}"] = res.numpy()                        # (N, 3), before +0.5 / clamp
    # build_rotation, quaternion_raw_multiply
    ns = lift(f"{REF}/lib/utils/general_utils.py", ["build_rotation", "quaternion_raw_multiply"])
    q = rng.standard_normal((N, 4)).astype(np.float32) * rng.uniform(0.5, 2.0, (N, 1)).astype(np.float32)
    q2 = rng.standard_normal((N, 4)).astype(np.float32)
    with cuda_to_cpu():
        out["quat"] = q
        out["rotmat"] = ns["build_rotation"](torch.from_numpy(q)).numpy()
        out["quat_b"] = q2
        out["quat_mul"] = ns["quaternion_raw_multiply"](None, torch.from_numpy(q), torch.from_numpy(q2)).numpy()
        # build2DRectangle
        ns2 = lift(f"{REF}/lib/utils/primitive_utils.py", ["build2DRectangle"], dict(build_rotation=ns["build_rotation"]))
        means = rng.uniform(-20, 20, (N, 3)).astype(np.float32)
        scales = np.exp(rng.normal(-2.0, 0.5, (N, 2))).astype(np.float32)
        opac = np.clip(1 / (1 + np.exp(-2 * rng.standard_normal((N, 1)))), 0.01, 0.999).astype(np.float32)
        v, f, _ = ns2["build2DRectangle"](torch.from_numpy(means), torch.from_numpy(scales), torch.from_numpy(q),
                                          torch.from_numpy(opac))
        out["rect_means"] = means; out["rect_scales"] = scales; out["rect_opac"] = opac
        out["rect_vertices"] = v.numpy(); out["rect_faces"] = f.numpy()
        # get_range_rays (Waymo: list of inclinations, offset .5; KITTI: 2 bounds, offset 0)
        ns3 = lift(f"{REF}/lib/scene/lidar_sensor.py", ["LiDARSensor.get_range_rays"])
        for tag, H, W, inc, off in [("waymo", 8, 16, syn.waymo_inclinations(8).tolist(), 0.5),
                                    ("kitti", 6, 12, [float(np.radians(-24.9)), float(np.radians(2.0))], 0.0)]:
            pose = torch.from_numpy(syn.sensor_pose(3))
            me = types.SimpleNamespace(inclination_bounds=inc, sensor2world={0: pose}, sensor_center={0: pose[:3, 3]},
                                       H=H, W=W, pixel_offset=off, angle_offset=0.0)
            ro, rd = ns3["get_range_rays"](me, 0)
            out[f"rays_{tag}_o"] = ro.contiguous().numpy(); out[f"rays_{tag}_d"] = rd.numpy()
            out[f"rays_{tag}_pose"] = pose.numpy(); out[f"rays_{tag}_inc"] = np.asarray(inc, np.float32)
    np.savez_compressed(os.path.join(OUT, "ref_python.npz"), **out)
    print("ref_python.npz", {k: v.shape for k, v in out.items()})


def lift_statements(path: str, func: str, first_target: str, last_target: str):
    """The statements of `func`'s body from the first assignment to `first_target` through the assignment to `last_target`,
    compiled as they stand (used to run the middle of raytracing() without its tracer calls)."""
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == func)
    tgt = lambda n: [t.id for t in n.targets if isinstance(t, ast.Name)] if isinstance(n, ast.Assign) else []
    i0 = next(i for i, n in enumerate(fn.body) if first_target in tgt(n))
    i1 = max(i for i, n in enumerate(fn.body) if last_target in tgt(n))
    return compile(ast.Module(fn.body[i0:i1 + 1], []), path, "exec"), (fn.body[i0].lineno, fn.body[i1].end_lineno)


def gen_prepare_fixtures():
    """what lrt_prepare / lrt_prepare_backward replace, produced by the reference's own statements."""
    rng = np.random.default_rng(17)
    gu = lift(f"{REF}/lib/utils/general_utils.py", ["build_rotation", "quaternion_raw_multiply"])
    gm = lift(f"{REF}/lib/scene/gaussian_model.py",
              ["GaussianModel.setup_functions", "GaussianModel.get_scaling", "GaussianModel.get_rotation", "GaussianModel.get_world_xyz",
               "GaussianModel.get_features", "GaussianModel.get_opacity"],
              dict(build_rotation=gu["build_rotation"], inverse_sigmoid=None), strip_decorators=True)
    body, lines = lift_statements(f"{REF}/lib/gaussian_renderer/__init__.py", "raytracing", "all_means3D", "colors_precomp")

    class Asset:                                   # the accessor surface raytracing() uses, bound to the lifted methods
        def __init__(self, P, pose):
            t = lambda a: torch.tensor(np.asarray(a, np.float32), requires_grad=True)
            self._xyz = t(rng.uniform(-3, 3, (P, 3))); self._scaling = t(rng.normal(-2.0, 0.6, (P, 2)))
            self._rotation = t(rng.standard_normal((P, 4)) * rng.uniform(0.3, 3.0, (P, 1)))
            self._opacity = t(rng.normal(0, 2, (P, 1)))
            self._features_dc = t(rng.standard_normal((P, 1, 3))); self._features_rest = t(0.1 * rng.standard_normal((P, 15, 3)))
            self.active_sh_degree = 3; self.max_sh_degree = 3
            self.bounding_box = None if pose is None else types.SimpleNamespace(frame={1: pose})
            gm["setup_functions"](self)
        get_scaling = property(gm["get_scaling"]); get_opacity = property(gm["get_opacity"]); get_features = property(gm["get_features"])
        def get_rotation(self, ts=0.0): return gm["get_rotation"](self, ts)
        def get_world_xyz(self, ts=0.0): return gm["get_world_xyz"](self, ts)
        def leaves(self): return [self._xyz, self._scaling, self._rotation, self._opacity, self._features_dc, self._features_rest]

    def pose():
        return (torch.tensor(rng.uniform(-20, 20, 3).astype(np.float32)), torch.tensor((rng.standard_normal((1, 4)) * 1.4).astype(np.float32)))

    with cuda_to_cpu():
        assets = [Asset(300, None), Asset(100, pose()), Asset(120, pose())]
    out = {"lines": np.array(lines, np.int32), "n_assets": np.int32(len(assets))}
    names = ("xyz", "scaling", "rotation", "opacity", "features_dc", "features_rest")
    for k, a in enumerate(assets):
        for nm, t in zip(names, a.leaves()):
            out[f"asset{k}/{nm}"] = t.detach().numpy()
        if a.bounding_box is not None:
            out[f"asset{k}/pose_T"] = a.bounding_box.frame[1][0].numpy(); out[f"asset{k}/pose_quat"] = a.bounding_box.frame[1][1].numpy()
    for tag, dynamic, decomp in (("static", False, False), ("dynamic", True, False), ("object", True, "object"), ("background", True, "background")):
        use = assets[:1] if decomp == "background" else (assets[1:] if decomp == "object" else assets)
        ns = dict(torch=torch, gaussian_assets=use, frame=1, decomp=decomp, override_color=None, scaling_modifier=1.0, sensor_center=None,
                  args=types.SimpleNamespace(dynamic=dynamic, pipe=types.SimpleNamespace(compute_cov3D_python=False, convert_SHs_python=False)),
                  quaternion_raw_multiply=gu["quaternion_raw_multiply"])
        for a in assets:
            for t in a.leaves():
                t.grad = None
        with cuda_to_cpu():
            exec(body, ns)
        res = [ns[n] for n in ("means3D", "opacity", "scales", "rotations", "shs")]
        if not dynamic and len(use) > 1:
            # the reference takes rot_in_local[0] only for a static scene (:117-118): fixtures use one asset there
            use = use[:1]; ns["gaussian_assets"] = use
            with cuda_to_cpu():
                exec(body, ns)
            res = [ns[n] for n in ("means3D", "opacity", "scales", "rotations", "shs")]
        ws = [torch.tensor(rng.standard_normal(tuple(t.shape)).astype(np.float32)) for t in res]
        sum((t * w).sum() for t, w in zip(res, ws)).backward()
        out[f"{tag}/assets"] = np.array([assets.index(a) for a in use], np.int32)
        for n, t, w in zip(("means3D", "opacity", "scales", "rotations", "shs"), res, ws):
            out[f"{tag}/{n}"] = t.detach().numpy(); out[f"{tag}/w_{n}"] = w.numpy()
        for a in use:
            k = assets.index(a)
            for nm, t in zip(names, a.leaves()):
                out[f"{tag}/g_asset{k}/{nm}"] = t.grad.numpy().copy()
    # pinhole rays of the Camera branch (lib/gaussian_renderer/__init__.py:31-41 -> graphics_utils.get_rays)
    gr_ = lift(f"{REF}/lib/utils/graphics_utils.py", ["get_rays"])
    W_, H_, fovx = 40, 24, 1.2
    focal = 0.5 * W_ / np.tan(0.5 * fovx)
    K = np.array([[focal, 0, 0.5 * W_], [0, focal, 0.5 * H_], [0, 0, 1]])
    c2w = torch.tensor(syn.sensor_pose(4).astype(np.float32))[:3, :4]
    with cuda_to_cpu():
        ro, rd = gr_["get_rays"](K, c2w)
    out["cam/K"] = K; out["cam/c2w"] = c2w.numpy(); out["cam/rays_o"] = ro.numpy(); out["cam/rays_d"] = rd.numpy()
    out["cam/whf"] = np.array([W_, H_, fovx])
    np.savez_compressed(os.path.join(OUT, "ref_prepare.npz"), **out)
    print("ref_prepare.npz: raytracing() lines", lines, {k: v.shape for k, v in out.items() if k.startswith("dynamic/")})


def gen_densify_fixtures():
    """the reference's restructuring methods (gaussian_model.py:235-407), lifted and run unmodified on the CPU over a
    seeded model with a real torch.optim.Adam(l, lr=0.0, eps=1e-15) that has taken two steps. The split's torch.normal samples are
    recorded on the way so the native kernels can be fed the same numbers."""
    torch.manual_seed(23)
    rng = np.random.default_rng(23)
    gu = lift(f"{REF}/lib/utils/general_utils.py", ["build_rotation"])
    meths = ["setup_functions", "get_scaling", "get_opacity", "get_local_xyz", "_prune_optimizer", "prune_points", "cat_tensors_to_optimizer",
             "densification_postfix", "densify_and_split", "densify_and_clone", "densify_and_prune"]
    recorded = []
    real_normal = torch.normal

    def normal_rec(*a, **k):
        r = real_normal(*a, **k); recorded.append(r.clone()); return r
    class _TorchShim:                                  # `torch` as the lifted methods see it: normal() recorded, cuda.empty_cache() a no-op
        normal = staticmethod(normal_rec)
        cuda = types.SimpleNamespace(empty_cache=lambda: None)

        def __getattr__(self, k):
            return getattr(torch, k)
    tshim = _TorchShim()
    gm = lift(f"{REF}/lib/scene/gaussian_model.py", [f"GaussianModel.{m}" for m in meths],
              dict(build_rotation=gu["build_rotation"], inverse_sigmoid=None, nn=torch.nn, torch=tshim, print=lambda *a, **k: None), strip_decorators=True)
    P = 400

    class M:
        pass
    for m in meths:
        if m in ("get_scaling", "get_opacity", "get_local_xyz"):
            setattr(M, m, property(gm[m]))
        else:
            setattr(M, m, gm[m])
    me = M()
    me.setup_functions()
    t = lambda a: torch.nn.Parameter(torch.tensor(np.asarray(a, np.float32)))
    me._xyz = t(rng.uniform(-5, 5, (P, 3))); me._features_dc = t(rng.standard_normal((P, 1, 3))); me._features_rest = t(0.1 * rng.standard_normal((P, 15, 3)))
    me._opacity = t(rng.normal(0, 2.5, (P, 1))); me._scaling = t(rng.normal(-2.2, 0.7, (P, 2))); me._rotation = t(rng.standard_normal((P, 4)) * 1.5)
    me.dimension = 2; me.extent = 10.0; me.densify_scale_threshold = 0.012; me.bounding_box = None
    me.max_radii2D = torch.zeros(P)
    me.xyz_gradient_accum = torch.tensor(rng.uniform(0, 2e-3, (P, 1)).astype(np.float32)); me.denom = torch.tensor(rng.integers(0, 4, (P, 1)).astype(np.float32))
    names = [("xyz", "_xyz"), ("f_dc", "_features_dc"), ("f_rest", "_features_rest"), ("opacity", "_opacity"), ("scaling", "_scaling"), ("rotation", "_rotation")]
    me.optimizer = torch.optim.Adam([{"params": [getattr(me, a)], "lr": 1e-3 * (k + 1), "name": n} for k, (n, a) in enumerate(names)], lr=0.0, eps=1e-15)
    for _ in range(2):
        for n, a in names:
            getattr(me, a).grad = torch.tensor(rng.standard_normal(tuple(getattr(me, a).shape)).astype(np.float32))
        me.optimizer.step()
    out = {}

    def snap(tag):
        for n, a in names:
            p = getattr(me, a); st = me.optimizer.state[p]
            out[f"{tag}/{n}"] = p.detach().numpy().copy(); out[f"{tag}/{n}/exp_avg"] = st["exp_avg"].numpy().copy()
            out[f"{tag}/{n}/exp_avg_sq"] = st["exp_avg_sq"].numpy().copy(); out[f"{tag}/{n}/step"] = np.float32(float(st["step"]))
        out[f"{tag}/xyz_gradient_accum"] = me.xyz_gradient_accum.numpy().copy(); out[f"{tag}/denom"] = me.denom.numpy().copy()
        out[f"{tag}/max_radii2D"] = me.max_radii2D.numpy().copy()
    snap("in")
    # 1. prune_points alone
    mask = torch.tensor(rng.random(P) < 0.3)
    out["prune/mask"] = mask.numpy()
    import copy
    keep_state = copy.deepcopy((me.__dict__))
    with cuda_to_cpu():
        me.prune_points(mask)
    snap("prune")
    # 2. the whole densify_and_prune on the ORIGINAL model
    me.__dict__.update(keep_state)
    opt = types.SimpleNamespace(densify_grad_threshold=3e-4, thresh_opa_prune=0.05, prune_size_threshold=0.5)
    out["opt"] = np.array([opt.densify_grad_threshold, opt.thresh_opa_prune, opt.prune_size_threshold, me.densify_scale_threshold, me.extent], np.float64)
    with cuda_to_cpu():
        res = me.densify_and_prune(opt, 0.005, 20)
    out["densify/counts"] = np.array(res, np.int64)
    out["densify/samples"] = recorded[0].detach().numpy()
    snap("densify")
    np.savez_compressed(os.path.join(OUT, "ref_densify.npz"), **out)
    print("ref_densify.npz: prune keeps", int((~mask).sum()), "of", P, "; densify_and_prune (clone, split, prune_scale, prune_opacity) =", res,
          "->", out["densify/xyz"].shape[0], "rows; samples", out["densify/samples"].shape)


# ------------------------------------------------------------------ tracer fixtures from oracle/_ref
BG = np.array([0.0, 0.0, 1.0], np.float32)      # train.py:104-106


def run_ref(ref: Ref, o, d, sc: dict, D: int, dL=None):
    args = (o, d, BG, sc["means"], sc["scales"], sc["rots"], sc["opac"], sc["shs"], D)
    f = ref.forward(*args)
    res = dict(out=f["out"], accum_w=f["accum_w"])
    if dL is not None:
        g = ref.backward(*args, f["out"], dL)
        res.update({f"g_{k}": v for k, v in g.items()})
    return res


def kat_cases():
    """Hand-built known-answer scenes; every case is (name, scene dict, ray_o (1,3), ray_d (R,3), D)."""
    cases = []
    ident = np.array([[1.0, 0, 0, 0]], np.float32)
    rng = np.random.default_rng(11)

    def sh_const(P, c0=(0.3, 0.1, -0.2)):
        s = np.zeros((P, 16, 3), np.float32)
        s[:, 0, :] = np.asarray(c0, np.float32) / np.float32(syn.SH_C0)   # colour = c0 + 0.5
        return s

    def stack(n, z0, dz, op, s=0.5, jitter=0.0):
        means = np.zeros((n, 3), np.float32); means[:, 2] = z0 + dz * np.arange(n)
        means[:, :2] = jitter * rng.standard_normal((n, 2))
        return dict(means=means, scales=np.full((n, 2), s, np.float32), rots=np.repeat(ident, n, 0),
                    opac=np.full((n, 1), op, np.float32), shs=sh_const(n))

    zray = np.array([[0, 0, 1.0]], np.float32)
    fan = np.stack([np.array([0.02 * i, -0.015 * i, 1.0]) for i in range(8)]).astype(np.float32)
    fan /= np.linalg.norm(fan, axis=1, keepdims=True)
    o0 = np.zeros((1, 3), np.float32)
    cases.append(("single_onaxis", stack(1, 5.0, 0.0, 0.7), o0, zray, 0))
    cases.append(("two_stacked", stack(2, 5.0, 1.0, 0.6), o0, fan, 0))
    for n in (15, 16, 17, 32, 33, 40):
        cases.append((f"stack_{n}", stack(n, 3.0, 0.1, 0.05, jitter=0.02), o0, fan, 3))
    cases.append(("terminate", stack(12, 2.0, 0.25, 0.95), o0, fan, 0))
    sc = stack(3, 4.0, 1.0, 0.5); sc["opac"][1, 0] = 0.003           # < 1/255: NaN proxy, never hit
    cases.append(("below_alpha_min", sc, o0, fan, 0))
    sc = stack(4, 0.05, 0.1, 0.5)                                     # hits at t = .05, .15 (< 0.2), .25, .35
    cases.append(("near_cutoff_0p2", sc, o0, fan, 0))
    # 16 hits then a 17th inside the STEP_EPSILON window, then one beyond it
    sc = stack(19, 3.0, 0.1, 0.05); sc["means"][16, 2] = sc["means"][15, 2] + 4e-6
    sc["means"][17, 2] = sc["means"][15, 2] + 5e-4
    cases.append(("epsilon_gap", sc, o0, zray, 0))
    # rotated / anisotropic / un-normalised quaternions, off-centre hits, SH degree 3, grazing rays
    P = 24
    means = rng.uniform(-1, 1, (P, 3)).astype(np.float32); means[:, 2] = rng.uniform(2, 9, P)
    rots = rng.standard_normal((P, 4)).astype(np.float32) * 1.7
    sc = dict(means=means, scales=np.exp(rng.normal(-0.5, 0.5, (P, 2))).astype(np.float32), rots=rots,
              opac=rng.uniform(0.05, 0.95, (P, 1)).astype(np.float32),
              shs=(0.3 * rng.standard_normal((P, 16, 3))).astype(np.float32))
    d = rng.standard_normal((48, 3)).astype(np.float32) * np.array([0.25, 0.25, 0.0], np.float32) + np.array([0, 0, 1], np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    for D in range(4):
        cases.append((f"random_deg{D}", sc, np.array([[0.1, -0.05, 0.0]], np.float32), d, D))
    return cases


def gen_tracer_fixtures():
    ref = Ref()
    rng = np.random.default_rng(5)
    # --- KATs
    out = {}
    names = []
    for name, sc, o, d, D in kat_cases():
        R = d.shape[0]
        dL = np.zeros((R, 9), np.float32); dL[:, :4] = rng.standard_normal((R, 4)); dL[:, 5:8] = 0.1 * rng.standard_normal((R, 3))
        res = run_ref(ref, o, d.reshape(1, R, 3), sc, D, dL)
        names.append(name)
        for k, v in sc.items():
            out[f"{name}/{k}"] = v
        out[f"{name}/ray_o"] = o; out[f"{name}/ray_d"] = d; out[f"{name}/D"] = np.int32(D); out[f"{name}/dL"] = dL
        for k, v in res.items():
            out[f"{name}/{k}"] = v
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "ref_kat.npz"), **out)
    print("ref_kat.npz", len(names), "cases")

    # --- small seeded scene, forward + backward (inputs stored)
    scn = syn.make_street_scene(1500, seed=3, extent=25.0, scale_mult=0.5)
    o, d = syn.ray_patch(24, 32)
    R = d.shape[0] * d.shape[1]
    dL = np.zeros((R, 9), np.float32); dL[:, :4] = rng.standard_normal((R, 4))
    sc = dict(means=scn.means, scales=scn.scales, rots=scn.rots, opac=scn.opac, shs=scn.shs)
    res = run_ref(ref, o, d, sc, 3, dL)
    np.savez_compressed(os.path.join(OUT, "ref_scene_small.npz"), ray_o=o, ray_d=d, dL=dL, D=np.int32(3), **sc, **res)
    print("ref_scene_small.npz", {k: v.shape for k, v in res.items()})

    # --- BASELINE config #1: 10k Gaussians, 64 x 64 rays, forward (inputs regenerated from the seed)
    scn = syn.make_street_scene(10000, seed=0)
    o, d = syn.ray_patch(64, 64)
    sc = dict(means=scn.means, scales=scn.scales, rots=scn.rots, opac=scn.opac, shs=scn.shs)
    res = run_ref(ref, o, d, sc, 3)
    chk = np.array([float(np.abs(v.astype(np.float64)).sum()) for v in (scn.means, scn.scales, scn.rots, scn.opac, scn.shs, d)])
    np.savez_compressed(os.path.join(OUT, "ref_cfg1_forward.npz"), seed=np.int32(0), P=np.int32(10000),
                        input_checksums=chk, out=res["out"], accum_w=res["accum_w"])
    print("ref_cfg1_forward.npz", res["out"].shape)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["python", "prepare", "densify", "tracer"]
    if "python" in which:
        gen_python_fixtures()
    if "prepare" in which:
        gen_prepare_fixtures()
    if "densify" in which:
        gen_densify_fixtures()
    if "tracer" in which:
        gen_tracer_fixtures()

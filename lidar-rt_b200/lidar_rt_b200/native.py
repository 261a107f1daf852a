"""ctypes binding of the C-ABI library (include/lidar_rt_b200.h -> csrc/liblidar_rt_b200.so).

This is the thin host layer the reference implements as a pybind11 torch extension
(/root/reference/submodules/diff-lidar-tracer/ext.cpp:17-22). PyTorch is used only for device
memory and streams; tensors are handed to the library as raw device pointers.

There is NO fallback: if the library is missing or fails, the call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LIDAR_RT_B200_LIB") or os.path.join(os.path.dirname(_HERE), "csrc", "liblidar_rt_b200.so")

LRT_FLAG_FIX_BG_GRAD = 1
NUM_CHANNELS = 9
DEFAULT_HIT_CAP = 256          # contributing hits recorded per ray for the backward replay (rays beyond it are re-traced)
OPT_FORWARD_KERNEL, OPT_RAY_GRID_WIDTH, OPT_VECTOR_ATOMICS, OPT_MORTON_BITS, OPT_BACKWARD_KERNEL, OPT_WAVEFRONT_SHADE, OPT_KERNEL_TIMING, OPT_SORT_RAYS, OPT_BEAM_CELL_PCT = 1, 2, 3, 4, 5, 6, 7, 8, 9
OPT_TRIANGLE_DEPTH, OPT_SPLIT_FUSED, OPT_SORT_KEY_BITS, OPT_BIN_CAP = 10, 11, 12, 13


class LrtError(RuntimeError):
    pass


class LrtInfo(ctypes.Structure):
    _fields_ = [("P", c_int32), ("levels", c_int32), ("nodes", c_int64), ("bytes_records", c_int64),
                ("bytes_nodes", c_int64), ("bytes_workspace", c_int64), ("builds", c_int64), ("refits", c_int64),
                ("kernel_launches", c_int32)]


class LrtAsset(ctypes.Structure):
    """lrt_asset of include/lidar_rt_b200.h: one GaussianModel's leaf tensors (+ optional pose, + gradient targets)."""
    _fields_ = [("P", c_int32), ("compose_rotation", c_int32),
                ("xyz", c_void_p), ("scaling", c_void_p), ("rotation", c_void_p), ("opacity", c_void_p),
                ("features_dc", c_void_p), ("features_rest", c_void_p), ("pose_T", c_void_p), ("pose_quat", c_void_p),
                ("d_xyz", c_void_p), ("d_scaling", c_void_p), ("d_rotation", c_void_p), ("d_opacity", c_void_p),
                ("d_features_dc", c_void_p), ("d_features_rest", c_void_p)]


class LrtAdamTensor(ctypes.Structure):
    """lrt_adam_tensor of include/lidar_rt_b200.h: one parameter tensor with its gradient and Adam state."""
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("n", c_int64), ("lr", c_float), ("step", c_int32)]


class LrtShPart(ctypes.Structure):
    """lrt_sh_part of include/lidar_rt_b200.h: one asset's SH leaves (and, for the backward, their gradient buffers)."""
    _fields_ = [("P", c_int32), ("reserved", c_int32), ("features_dc", c_void_p), ("features_rest", c_void_p),
                ("d_features_dc", c_void_p), ("d_features_rest", c_void_p)]


class LrtRowTensor(ctypes.Structure):
    """lrt_row_tensor of include/lidar_rt_b200.h: one per-Gaussian tensor moved by lrt_compact_rows / lrt_densify_rows."""
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("row_floats", c_int32), ("kind", c_int32)]


ROW_COPY, ROW_ZERO_NEW, ROW_XYZ, ROW_SCALING = 0, 1, 2, 3
MAX_ROW_TENSORS = 32
MAX_ASSETS = 128
_lib = None


def load_library() -> ctypes.CDLL:
    """Load liblidar_rt_b200.so; raises if it has not been built (csrc/build.sh)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LrtError(f"{LIB_PATH} not found: build it with lidar-rt_b200/csrc/build.sh "
                       "(or __graft_entry__.build()); there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    fp, ip = c_void_p, c_void_p
    lib.lrt_version.restype = c_int
    lib.lrt_ctx_create.argtypes = [c_int, POINTER(c_void_p)]
    lib.lrt_ctx_destroy.argtypes = [c_void_p]
    lib.lrt_last_error.argtypes = [c_void_p]; lib.lrt_last_error.restype = c_char_p
    build_args = [c_void_p, c_int, fp, fp, fp, fp, c_float, c_void_p]
    lib.lrt_build.argtypes = build_args
    lib.lrt_refit.argtypes = build_args
    lib.lrt_forward.argtypes = [c_void_p, c_int, fp, c_int, fp, fp, c_int, fp, fp, fp, fp, fp, c_int, c_int, c_float,
                                fp, fp, ip, fp, fp, ip, c_int, ip, c_void_p]
    lib.lrt_backward.argtypes = [c_void_p, c_int, fp, c_int, fp, fp, c_int, fp, fp, fp, fp, fp, c_int, c_int, c_float,
                                 fp, fp, ip, fp, fp, ip, c_int, fp, fp, fp, fp, fp, c_int, c_void_p]
    lib.lrt_prepare.argtypes = [c_void_p, c_int, POINTER(LrtAsset), c_int, fp, fp, fp, fp, fp, c_void_p]
    lib.lrt_prepare_backward.argtypes = [c_void_p, c_int, POINTER(LrtAsset), c_int, fp, fp, fp, fp, fp, c_void_p]
    lib.lrt_prepare.restype = c_int; lib.lrt_prepare_backward.restype = c_int
    lib.lrt_range_rays.argtypes = [c_void_p, c_int, c_int, fp, c_float, c_float, c_float, c_float, fp, fp, fp, c_void_p]
    lib.lrt_range_points.argtypes = [c_void_p, c_int, c_int, fp, c_float, c_float, c_float, c_float, fp, fp, fp, c_void_p]
    lib.lrt_range_rays.restype = c_int; lib.lrt_range_points.restype = c_int
    lib.lrt_chamfer_forward.argtypes = [c_void_p, c_int, c_int, fp, c_int, fp, fp, ip, fp, ip, c_void_p]
    lib.lrt_chamfer_backward.argtypes = [c_void_p, c_int, c_int, fp, c_int, fp, fp, fp, ip, ip, fp, fp, c_void_p]
    lib.lrt_chamfer_forward.restype = c_int; lib.lrt_chamfer_backward.restype = c_int
    lib.lrt_adam_step.argtypes = [c_void_p, c_int, POINTER(LrtAdamTensor), c_double, c_double, c_double, c_void_p]
    lib.lrt_adam_step.restype = c_int
    lib.lrt_set_sh_parts.argtypes = [c_void_p, c_int, POINTER(LrtShPart), c_int, c_void_p]
    lib.lrt_set_sh_parts.restype = c_int
    lib.lrt_compact_rows.argtypes = [c_void_p, c_int, c_void_p, c_int, POINTER(LrtRowTensor), c_void_p]
    lib.lrt_densify_rows.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, fp, fp, c_int, POINTER(LrtRowTensor), c_void_p]
    lib.lrt_compact_rows.restype = c_int; lib.lrt_densify_rows.restype = c_int
    lib.lrt_set_option.argtypes = [c_void_p, c_int, c_int]
    lib.lrt_get_kernel_times.argtypes = [c_void_p, c_char_p, POINTER(c_float), POINTER(c_int), c_int]
    lib.lrt_get_kernel_times.restype = c_int
    lib.lrt_get_info.argtypes = [c_void_p, POINTER(LrtInfo)]
    lib.lrt_get_permutation.argtypes = [c_void_p, c_void_p, c_void_p]
    for f in (lib.lrt_set_option, lib.lrt_ctx_create, lib.lrt_ctx_destroy, lib.lrt_build, lib.lrt_refit, lib.lrt_forward, lib.lrt_backward,
              lib.lrt_get_info, lib.lrt_get_permutation):
        f.restype = c_int
    _lib = lib
    return lib


def _f32(t: torch.Tensor, name: str, shape_tail=None, device=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise LrtError(f"{name} must be a CUDA tensor")
    if device is not None and t.device != device:
        # raw pointers of another GPU would reach kernels launched on the context's device (illegal address or silent peer reads)
        raise LrtError(f"{name} lives on {t.device}, the context on {device}")
    if t.dtype != torch.float32:
        raise LrtError(f"{name} must be float32, got {t.dtype}")
    if shape_tail is not None and tuple(t.shape[-len(shape_tail):]) != tuple(shape_tail):
        raise LrtError(f"{name} must have trailing shape {shape_tail}, got {tuple(t.shape)}")
    return t.contiguous()


def _ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def _stream(device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Context:
    """Owns one lrt_ctx (acceleration structure + workspace) on one device."""

    def __init__(self, device=None):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise LrtError("lidar_rt_b200 needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else torch.device(device).index or 0)
        h = c_void_p()
        rc = self.lib.lrt_ctx_create(self.device.index, byref(h))
        if rc != 0:
            raise LrtError(self.lib.lrt_last_error(None).decode())
        self._h = h
        self.generation = 0          # bumps on every build / refit
        # tuning knobs for A/B runs without code changes: LIDAR_RT_B200_OPTIONS="8=0,5=1" (LRT_OPT_* id = value)
        for kv in filter(None, os.environ.get("LIDAR_RT_B200_OPTIONS", "").split(",")):
            k, v = kv.split("=")
            self.set_option(int(k), int(v))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.lrt_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise LrtError(self.lib.lrt_last_error(self._h).decode())

    # ---- acceleration structure
    def _gauss(self, means, scales, rots, opac):
        means = _f32(means, "means3D", (3,), self.device)
        P = means.shape[0]
        if means.dim() != 2:
            raise LrtError("means3D must have dimensions (num_points, 3)")      # trace_surfels.cpp:178-180
        scales = _f32(scales, "scales", (2,), self.device); rots = _f32(rots, "rotations", (4,), self.device)
        opac = _f32(opac, "opacities", None, self.device).reshape(-1)
        if scales.shape[0] != P or rots.shape[0] != P or opac.shape[0] != P:
            raise LrtError("means3D / scales / rotations / opacities disagree on the number of Gaussians")
        return P, means, scales, rots, opac

    def build(self, means, scales, rots, opac, scale_modifier: float = 1.0, refit: bool = False):
        P, means, scales, rots, opac = self._gauss(means, scales, rots, opac)
        with torch.cuda.device(self.device):
            fn = self.lib.lrt_refit if refit else self.lib.lrt_build
            self._check(fn(self._h, P, _ptr(means), _ptr(scales), _ptr(rots), _ptr(opac), c_float(scale_modifier),
                           _stream(self.device)))
        self.generation += 1
        return self.generation

    def set_option(self, option: int, value: int):
        self._check(self.lib.lrt_set_option(self._h, int(option), int(value)))

    def kernel_times(self) -> dict:
        """{kernel name: (total ms, launches)} of the CUDA-event spans recorded since the previous call
        (set_option(OPT_KERNEL_TIMING, 1) first). Synchronises on the recorded events."""
        cap = 32
        names = ctypes.create_string_buffer(32 * cap)
        ms = (c_float * cap)(); cnt = (c_int * cap)()
        n = self.lib.lrt_get_kernel_times(self._h, names, ms, cnt, cap)
        if n < 0:
            raise LrtError("lrt_get_kernel_times failed")
        return {names.raw[32 * i:32 * (i + 1)].split(b"\0")[0].decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def info(self) -> LrtInfo:
        i = LrtInfo()
        self._check(self.lib.lrt_get_info(self._h, byref(i)))
        return i

    def permutation(self) -> torch.Tensor:
        P = self.info().P
        out = torch.empty(P, dtype=torch.int32, device=self.device)
        self._check(self.lib.lrt_get_permutation(self._h, _ptr(out), _stream(self.device)))
        return out

    # ---- fused activation + world transform + concatenation
    @staticmethod
    def _asset_table(assets, grads=None):
        """assets: list of dicts with the leaf tensors xyz, scaling, rotation, opacity, features_dc, features_rest (float32 CUDA,
        contiguous), optional pose_T (3,), pose_quat (4,), compose_rotation. grads: matching list of dicts of output tensors."""
        n = len(assets)
        if not 1 <= n <= MAX_ASSETS:
            raise LrtError(f"need 1..{MAX_ASSETS} assets, got {n}")
        tab = (LrtAsset * n)()
        M = None
        for k, a in enumerate(assets):
            leaves = {}
            for name in ("xyz", "scaling", "rotation", "opacity", "features_dc", "features_rest"):
                t = a[name]
                if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                    raise LrtError(f"asset {k}: {name} must be a contiguous float32 CUDA tensor")
                leaves[name] = t
            P = leaves["xyz"].shape[0]
            if (tuple(leaves["xyz"].shape) != (P, 3) or tuple(leaves["scaling"].shape) != (P, 2) or tuple(leaves["rotation"].shape) != (P, 4)
                    or leaves["opacity"].numel() != P or tuple(leaves["features_dc"].shape) != (P, 1, 3)
                    or leaves["features_rest"].shape[0] != P or leaves["features_rest"].shape[2:] != (3,)):
                raise LrtError(f"asset {k}: leaf tensors disagree on shapes")
            m_k = 1 + leaves["features_rest"].shape[1]
            if M is None:
                M = m_k
            elif M != m_k:
                raise LrtError("assets disagree on the number of SH coefficients")
            e = tab[k]
            e.P = P; e.compose_rotation = int(bool(a.get("compose_rotation", False)))
            for name in leaves:
                setattr(e, name, leaves[name].data_ptr())
            for name in ("pose_T", "pose_quat"):
                t = a.get(name)
                if t is not None:
                    if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                        raise LrtError(f"asset {k}: {name} must be a contiguous float32 CUDA tensor")
                    setattr(e, name, t.data_ptr())
            if grads is not None:
                for name, t in grads[k].items():
                    if t is not None:
                        setattr(e, "d_" + name, t.data_ptr())
        return tab, n, M

    def prepare(self, assets, with_shs: bool = True):
        """lrt_prepare: -> (means (P,3), scales (P,2), rots (P,4), opac (P,1), shs (P,M,3)) for the concatenated assets.
        with_shs=False skips the concatenated SH copy (shs is None: the tracer reads the leaves in place, set_sh_parts)."""
        tab, n, M = self._asset_table(assets)
        P = sum(int(tab[k].P) for k in range(n))
        dev = self.device
        with torch.cuda.device(dev):
            means = torch.empty((P, 3), dtype=torch.float32, device=dev); scales = torch.empty((P, 2), dtype=torch.float32, device=dev)
            rots = torch.empty((P, 4), dtype=torch.float32, device=dev); opac = torch.empty((P, 1), dtype=torch.float32, device=dev)
            shs = torch.empty((P, M, 3), dtype=torch.float32, device=dev) if with_shs else None
            self._check(self.lib.lrt_prepare(self._h, n, tab, M, _ptr(means), _ptr(scales), _ptr(rots), _ptr(opac), _ptr(shs), _stream(dev)))
        return means, scales, rots, opac, shs

    def prepare_backward(self, assets, g_means, g_scales, g_rots, g_opac, g_shs, want=None):
        """lrt_prepare_backward: leaf gradients, one dict per asset (keys like the leaves). g_shs=None: the SH leaf gradients were
        written in place by the tracer's backward (they are not produced here)."""
        names = ("xyz", "scaling", "rotation", "opacity", "features_dc", "features_rest") if g_shs is not None else ("xyz", "scaling", "rotation", "opacity")
        dev = self.device
        with torch.cuda.device(dev):
            # one allocation per leaf kind, carved into per-asset views (6 allocations instead of 6 per asset)
            grads = [dict.fromkeys(names) for _ in assets]
            for nm in names:
                ks = [k for k in range(len(assets)) if want is None or want[k].get(nm, True)]
                if not ks:
                    continue
                sizes = [assets[k][nm].numel() for k in ks]
                flat = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
                for k, part in zip(ks, flat.split(sizes)):
                    grads[k][nm] = part.view(assets[k][nm].shape)
            tab, n, M = self._asset_table(assets, grads)
            args = [None if g is None else _f32(g, nm) for g, nm in ((g_means, "dL_dmeans"), (g_scales, "dL_dscales"), (g_rots, "dL_drots"), (g_opac, "dL_dopac"), (g_shs, "dL_dshs"))]
            self._check(self.lib.lrt_prepare_backward(self._h, n, tab, M, *(_ptr(t) for t in args), _stream(dev)))
        return grads

    # ---- range-image ray generation / back-projection
    def _grid(self, H, W, inclinations, sensor2world):
        dev = self.device
        s2w = torch.as_tensor(sensor2world, dtype=torch.float32, device=dev).reshape(4, 4).contiguous()
        table, lo, hi = None, 0.0, 0.0
        if isinstance(inclinations, torch.Tensor) or len(inclinations) != 2:
            table = torch.as_tensor(inclinations, dtype=torch.float32, device=dev).reshape(-1).contiguous()
            if table.numel() != H:
                raise LrtError(f"need H = {H} inclinations, got {table.numel()}")
        else:
            lo, hi = float(inclinations[0]), float(inclinations[1])
        return s2w, table, lo, hi

    def range_rays(self, H, W, inclinations, sensor2world, pixel_offset=0.5, angle_offset=0.0):
        """-> (rays_o (H,W,3) stride-0 view of the sensor centre, rays_d (H,W,3)): LiDARSensor.get_range_rays."""
        s2w, table, lo, hi = self._grid(H, W, inclinations, sensor2world)
        dev = self.device
        with torch.cuda.device(dev):
            d = torch.empty((H, W, 3), dtype=torch.float32, device=dev); c = torch.empty(3, dtype=torch.float32, device=dev)
            self._check(self.lib.lrt_range_rays(self._h, H, W, _ptr(table), c_float(lo), c_float(hi), c_float(pixel_offset),
                                                c_float(angle_offset), _ptr(s2w), _ptr(d), _ptr(c), _stream(dev)))
        return c[None, None].expand(H, W, 3), d

    def range_points(self, range_map, inclinations, sensor2world, pixel_offset=0.5, angle_offset=0.0):
        """-> world points (H,W,3) of a range map (H,W): LiDARSensor.range2point."""
        rm = _f32(range_map, "range_map")
        if rm.dim() != 2:
            raise LrtError("range_map must be (H, W)")
        H, W = rm.shape
        s2w, table, lo, hi = self._grid(H, W, inclinations, sensor2world)
        dev = self.device
        with torch.cuda.device(dev):
            p = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
            self._check(self.lib.lrt_range_points(self._h, H, W, _ptr(table), c_float(lo), c_float(hi), c_float(pixel_offset),
                                                  c_float(angle_offset), _ptr(s2w), _ptr(rm), _ptr(p), _stream(dev)))
        return p

    # ---- Adam step over many tensors
    def adam_step(self, rows, beta1: float, beta2: float, eps: float):
        """lrt_adam_step. rows: iterable of (param, grad, exp_avg, exp_avg_sq, lr, step); all float32 CUDA contiguous, same numel."""
        rows = list(rows)
        tab = (LrtAdamTensor * max(len(rows), 1))()
        for k, (p, g, m, v, lr, step) in enumerate(rows):
            for t, nm in ((p, "param"), (g, "grad"), (m, "exp_avg"), (v, "exp_avg_sq")):
                if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                    raise LrtError(f"adam_step: row {k}: {nm} must be a contiguous float32 CUDA tensor")
                if t.device != self.device:
                    raise LrtError(f"adam_step: row {k}: {nm} lives on {t.device}, the context on {self.device}")
            n = p.numel()
            if g.numel() != n or m.numel() != n or v.numel() != n:
                raise LrtError(f"adam_step: row {k}: param / grad / state sizes differ")
            e = tab[k]
            e.param, e.grad, e.exp_avg, e.exp_avg_sq = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
            e.n, e.lr, e.step = n, float(lr), int(step)
        with torch.cuda.device(self.device):
            self._check(self.lib.lrt_adam_step(self._h, len(rows), tab, c_double(beta1), c_double(beta2), c_double(eps), _stream(self.device)))

    def adam_step_table(self, table, beta1: float, beta2: float, eps: float):
        """lrt_adam_step on a prepared table: a C-contiguous numpy array whose 48-byte rows are laid out like lrt_adam_tensor
        (optim.FusedAdam keeps one between steps). The caller vouches for the pointers."""
        if table.dtype.itemsize != ctypes.sizeof(LrtAdamTensor) or not table.flags["C_CONTIGUOUS"]:
            raise LrtError("adam_step_table: rows must be contiguous lrt_adam_tensor records")
        with torch.cuda.device(self.device):
            self._check(self.lib.lrt_adam_step(self._h, int(table.shape[0]), ctypes.cast(table.ctypes.data, POINTER(LrtAdamTensor)),
                                               c_double(beta1), c_double(beta2), c_double(eps), _stream(self.device)))

    # ---- densify / prune row compaction
    def _row_table(self, rows, n_in):
        rows = list(rows)
        if not 1 <= len(rows) <= MAX_ROW_TENSORS:
            raise LrtError(f"need 1..{MAX_ROW_TENSORS} row tensors, got {len(rows)}")
        tab = (LrtRowTensor * len(rows))()
        for k, (src, dst, kind) in enumerate(rows):
            for t, nm in ((src, "src"), (dst, "dst")):
                if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device:
                    raise LrtError(f"row tensor {k}: {nm} must be a contiguous float32 CUDA tensor on {self.device}")
            if src.shape[0] != n_in or tuple(src.shape[1:]) != tuple(dst.shape[1:]):
                raise LrtError(f"row tensor {k}: src must have {n_in} rows and dst the same row shape")
            rf = 1
            for d in src.shape[1:]:
                rf *= int(d)
            tab[k].src, tab[k].dst, tab[k].row_floats, tab[k].kind = src.data_ptr(), dst.data_ptr(), max(rf, 1), int(kind)
        return tab, len(rows), rows

    @staticmethod
    def _mask(m, n, name):
        if not isinstance(m, torch.Tensor) or not m.is_cuda or m.numel() != n:
            raise LrtError(f"{name} must be a CUDA tensor with {n} entries")
        return (m.reshape(-1) != 0).to(torch.uint8).contiguous()

    def compact_rows(self, keep, rows):
        """lrt_compact_rows: dst = src[keep] for every (src, dst, kind) of `rows` (kind ignored). dst must have keep.sum() rows."""
        rows = list(rows)
        n = rows[0][0].shape[0]
        k8 = self._mask(keep, n, "keep")
        if n == 0 or all(dst.shape[0] == 0 for _, dst, _ in rows):
            return                                   # nothing survives: the (empty) outputs are already what src[keep] is
        tab, nt, _ = self._row_table(rows, n)
        with torch.cuda.device(self.device):
            self._check(self.lib.lrt_compact_rows(self._h, n, _ptr(k8), nt, tab, _stream(self.device)))

    def densify_rows(self, clone_mask, split_mask, n_clone, n_split, N, samples, rotation, rows):
        """lrt_densify_rows: clone + split + removal of the split parents in one pass (see include/lidar_rt_b200.h)."""
        rows = list(rows)
        P = rows[0][0].shape[0]
        c8, s8 = self._mask(clone_mask, P, "clone_mask"), self._mask(split_mask, P, "split_mask")
        n_out = P - n_split + n_clone + N * n_split
        for k, (_, dst, _) in enumerate(rows):
            if dst.shape[0] != n_out:
                raise LrtError(f"row tensor {k}: dst must have {n_out} rows")
        smp = rot = None
        if n_split > 0:
            smp = _f32(samples, "samples", (3,), self.device)
            if smp.shape[0] != N * n_split:
                raise LrtError("samples must be (N * n_split, 3)")
            rot = _f32(rotation, "rotation", (4,), self.device)
        tab, nt, _ = self._row_table(rows, P)
        with torch.cuda.device(self.device):
            self._check(self.lib.lrt_densify_rows(self._h, P, _ptr(c8), _ptr(s8), int(n_clone), int(n_split), int(N), _ptr(smp), _ptr(rot), nt, tab,
                                                  _stream(self.device)))

    # ---- Chamfer distance
    @staticmethod
    def _clouds(xyz1, xyz2):
        a = _f32(xyz1, "xyz1", (3,)); c = _f32(xyz2, "xyz2", (3,))
        if a.dim() != 3 or c.dim() != 3 or a.shape[0] != c.shape[0]:
            raise LrtError("xyz1 / xyz2 must be (batch, n, 3) and (batch, m, 3)")
        return a, c, a.shape[0], a.shape[1], c.shape[1]

    def chamfer_forward(self, xyz1, xyz2):
        """lrt_chamfer_forward: -> (dist1 (b,n), dist2 (b,m), idx1 (b,n) int32, idx2 (b,m) int32), chamfer_3D.forward's outputs."""
        a, c, b, n, m = self._clouds(xyz1, xyz2)
        dev = self.device
        with torch.cuda.device(dev):
            d1 = torch.empty((b, n), dtype=torch.float32, device=dev); d2 = torch.empty((b, m), dtype=torch.float32, device=dev)
            i1 = torch.empty((b, n), dtype=torch.int32, device=dev); i2 = torch.empty((b, m), dtype=torch.int32, device=dev)
            self._check(self.lib.lrt_chamfer_forward(self._h, b, n, _ptr(a), m, _ptr(c), _ptr(d1), _ptr(i1), _ptr(d2), _ptr(i2), _stream(dev)))
        return d1, d2, i1, i2

    def chamfer_backward(self, xyz1, xyz2, grad_dist1, grad_dist2, idx1, idx2):
        """lrt_chamfer_backward: -> (grad_xyz1 (b,n,3), grad_xyz2 (b,m,3))."""
        a, c, b, n, m = self._clouds(xyz1, xyz2)
        g1 = _f32(grad_dist1, "grad_dist1"); g2 = _f32(grad_dist2, "grad_dist2")
        if tuple(g1.shape) != (b, n) or tuple(g2.shape) != (b, m) or tuple(idx1.shape) != (b, n) or tuple(idx2.shape) != (b, m):
            raise LrtError("grad_dist / idx shapes must be (batch, n) and (batch, m)")
        if idx1.dtype != torch.int32 or idx2.dtype != torch.int32 or not idx1.is_cuda or not idx2.is_cuda:
            raise LrtError("idx1 / idx2 must be int32 CUDA tensors")
        dev = self.device
        with torch.cuda.device(dev):
            ga = torch.empty_like(a); gc = torch.empty_like(c)
            self._check(self.lib.lrt_chamfer_backward(self._h, b, n, _ptr(a), m, _ptr(c), _ptr(g1), _ptr(g2), _ptr(idx1.contiguous()),
                                                      _ptr(idx2.contiguous()), _ptr(ga), _ptr(gc), _stream(dev)))
        return ga, gc

    # ---- SH rows in place
    def set_sh_parts(self, parts, grads=None):
        """lrt_set_sh_parts. parts: list of (features_dc (P,1,3), features_rest (P,M-1,3)) float32 CUDA contiguous tensors, in
        concatenation order; grads: matching list of (d_features_dc, d_features_rest) tensors (or None entries) for lrt_backward.
        Returns (P_total, M). An empty list unbinds."""
        n = len(parts)
        tab = (LrtShPart * max(n, 1))()
        total, M = 0, None
        for k, (dc, rest) in enumerate(parts):
            for t, nm in ((dc, "features_dc"), (rest, "features_rest")):
                if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device:
                    raise LrtError(f"SH part {k}: {nm} must be a contiguous float32 CUDA tensor on {self.device}")
            P = dc.shape[0]
            if tuple(dc.shape) != (P, 1, 3) or rest.dim() != 3 or rest.shape[0] != P or rest.shape[2] != 3:
                raise LrtError(f"SH part {k}: need features_dc (P,1,3) and features_rest (P,M-1,3)")
            m_k = 1 + rest.shape[1]
            if M is None:
                M = m_k
            elif M != m_k:
                raise LrtError("SH parts disagree on the number of coefficients")
            tab[k].P = P; tab[k].features_dc = dc.data_ptr(); tab[k].features_rest = rest.data_ptr()
            if grads is not None:
                gdc, grest = grads[k]
                for t, ref in ((gdc, dc), (grest, rest)):
                    if t is not None and (not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != ref.numel()):
                        raise LrtError(f"SH part {k}: gradient buffers must be contiguous float32 CUDA tensors shaped like the leaves")
                tab[k].d_features_dc = None if gdc is None else gdc.data_ptr()
                tab[k].d_features_rest = None if grest is None else grest.data_ptr()
            total += P
        with torch.cuda.device(self.device):
            self._check(self.lib.lrt_set_sh_parts(self._h, n, tab, int(M or 1), _stream(self.device)))
        return total, M

    # ---- one frame of the hot path as a CUDA graph
    def graphed_step(self, ray_o, ray_d, dL_dout, bg, means, scales, rots, opac, shs, sh_degree: int, scale_modifier: float = 1.0,
                     refit: bool = False, backward: bool = True, cap: int = DEFAULT_HIT_CAP):
        """Capture (re)build + forward (+ backward) of one frame into a CUDA graph and return a GraphedStep. The ~45 launches of
        a step cost the host ~0.6 ms to enqueue against ~2 ms of device time; replaying the captured graph costs one launch.
        The Gaussian tensors are read in place at every replay (update them in place between replays: an optimiser step does);
        rays and upstream gradients are copied into the graph's static buffers by GraphedStep.run()."""
        return GraphedStep(self, ray_o, ray_d, dL_dout, bg, means, scales, rots, opac, shs, sh_degree, scale_modifier, refit, backward, cap)

    # ---- forward / backward
    def _rays(self, ray_o, ray_d):
        ray_d = _f32(ray_d, "ray_d", (3,), self.device)
        lead = tuple(ray_d.shape[:-1])
        R = ray_d.numel() // 3
        if not isinstance(ray_o, torch.Tensor) or ray_o.dtype != torch.float32 or not ray_o.is_cuda:
            raise LrtError("ray_o must be a float32 CUDA tensor")
        if ray_o.device != self.device:
            raise LrtError(f"ray_o lives on {ray_o.device}, the context on {self.device}")
        # LiDARSensor.get_range_rays returns one centre .expand()-ed to (H,W,3): a stride-0 view.
        # The reference calls .contiguous() on a temporary (trace_surfels.cpp:227); we pass stride 0.
        if ray_o.numel() == 3 or (ray_o.shape[-1] == 3 and all(s == 0 for s in ray_o.stride()[:-1]) and ray_o.stride(-1) == 1):
            o = ray_o.reshape(-1)[:3] if ray_o.numel() == 3 else ray_o[(0,) * (ray_o.dim() - 1)]
            return R, lead, o.contiguous(), 0, ray_d
        if tuple(ray_o.shape) != tuple(ray_d.shape):
            raise LrtError("ray_o and ray_d must have the same shape")
        return R, lead, ray_o.contiguous(), 3, ray_d

    def forward(self, ray_o, ray_d, bg, means, scales, rots, opac, shs, sh_degree: int, scale_modifier: float = 1.0,
                record_hits: bool = True, cap: int = DEFAULT_HIT_CAP, want_slots: bool = False, record_aux: bool = True, sh_M: int = 16):
        P, means, scales, rots, opac = self._gauss(means, scales, rots, opac)
        R, lead, o, stride, d = self._rays(ray_o, ray_d)
        if shs is None:                  # rows read in place: set_sh_parts() was called for these Gaussians
            M = int(sh_M)
        else:
            shs = _f32(shs, "shs", (3,), self.device)
            if shs.dim() != 3 or shs.shape[0] != P:
                raise LrtError("shs must have dimensions (num_points, M, 3)")
            M = shs.shape[1]
        bg = _f32(bg, "bg", None, self.device).reshape(-1)
        dev = self.device
        self.set_option(OPT_RAY_GRID_WIDTH, lead[-1] if len(lead) >= 2 else 0)     # (H, W, 3) range image -> 4 x 8 warp tiles
        with torch.cuda.device(dev):
            out = torch.empty(lead + (NUM_CHANNELS,), dtype=torch.float32, device=dev)
            accum = torch.empty(P, dtype=torch.float32, device=dev)
            hit_g = hit_t = hit_a = hit_c = slots = None
            if record_hits:
                hit_g = torch.empty((cap, R), dtype=torch.int32, device=dev)
                hit_t = torch.empty((cap, R), dtype=torch.float32, device=dev)
                hit_c = torch.empty(R, dtype=torch.int32, device=dev)
                if record_aux:
                    hit_a = torch.empty((cap, R, 4), dtype=torch.float32, device=dev)    # (alpha, c0, c1, c2) per recorded hit
            if want_slots:
                slots = torch.empty(R, dtype=torch.int32, device=dev)
            self._check(self.lib.lrt_forward(self._h, R, _ptr(o), stride, _ptr(d), _ptr(bg), P, _ptr(means), _ptr(scales),
                                             _ptr(rots), _ptr(opac), _ptr(shs), int(sh_degree), M, c_float(scale_modifier),
                                             _ptr(out), _ptr(accum), _ptr(hit_g), _ptr(hit_t), _ptr(hit_a), _ptr(hit_c), cap, _ptr(slots),
                                             _stream(dev)))
        return dict(out=out, accum_w=accum, hit_gidx=hit_g, hit_t=hit_t, hit_aux=hit_a, hit_cnt=hit_c, slot_cnt=slots, cap=cap)

    def backward(self, ray_o, ray_d, bg, means, scales, rots, opac, shs, sh_degree: int, fwd_out, dL_dout,
                 hits: dict | None = None, scale_modifier: float = 1.0, flags: int = 0, sh_M: int = 16):
        """shs=None: SH rows and their gradients in place (set_sh_parts with gradient buffers first); the returned dict then has no "shs"."""
        P, means, scales, rots, opac = self._gauss(means, scales, rots, opac)
        R, lead, o, stride, d = self._rays(ray_o, ray_d)
        if shs is None:
            M = int(sh_M)
        else:
            shs = _f32(shs, "shs", (3,), self.device); M = shs.shape[1]
        bg = _f32(bg, "bg", None, self.device).reshape(-1)
        fwd_out = _f32(fwd_out, "out_attr_float32", (NUM_CHANNELS,), self.device); dL = _f32(dL_dout, "dL_dout", (NUM_CHANNELS,), self.device)
        if fwd_out.numel() != R * NUM_CHANNELS or dL.numel() != R * NUM_CHANNELS:
            raise LrtError("out / dL_dout must be (..., 9) matching the rays")
        dev = self.device
        self.set_option(OPT_RAY_GRID_WIDTH, lead[-1] if len(lead) >= 2 else 0)
        with torch.cuda.device(dev):
            g_means = torch.empty((P, 3), dtype=torch.float32, device=dev)
            g_shs = torch.empty((P, M, 3), dtype=torch.float32, device=dev) if shs is not None else None
            g_opac = torch.empty((P, 1), dtype=torch.float32, device=dev)
            g_scales = torch.empty((P, 2), dtype=torch.float32, device=dev)
            g_rots = torch.empty((P, 4), dtype=torch.float32, device=dev)
            hg = ht = ha = hc = None; cap = 0
            if hits is not None and hits.get("hit_gidx") is not None:
                hg, ht, hc, cap = hits["hit_gidx"], hits["hit_t"], hits["hit_cnt"], int(hits["cap"])
                ha = hits.get("hit_aux")
            self._check(self.lib.lrt_backward(self._h, R, _ptr(o), stride, _ptr(d), _ptr(bg), P, _ptr(means), _ptr(scales),
                                              _ptr(rots), _ptr(opac), _ptr(shs), int(sh_degree), M, c_float(scale_modifier),
                                              _ptr(fwd_out), _ptr(dL), _ptr(hg), _ptr(ht), _ptr(ha), _ptr(hc), cap,
                                              _ptr(g_means), _ptr(g_shs), _ptr(g_opac), _ptr(g_scales), _ptr(g_rots),
                                              int(flags), _stream(dev)))
        res = dict(means=g_means, opac=g_opac, scales=g_scales, rots=g_rots)
        if g_shs is not None:
            res["shs"] = g_shs
        return res


class GraphedStep:
    """build / refit + lrt_forward (+ lrt_backward) of one frame captured once (torch.cuda.CUDAGraph on the calling stream) and
    replayed per frame. Workspace growth (cudaMalloc) cannot happen inside a capture, so the step runs eagerly twice first."""

    def __init__(self, ctx: Context, ray_o, ray_d, dL_dout, bg, means, scales, rots, opac, shs, sh_degree, scale_modifier, refit, backward, cap):
        self.ctx = ctx
        self.ray_o = ray_o.detach().reshape(-1)[:3].clone() if ray_o.numel() == 3 or all(s == 0 for s in ray_o.stride()[:-1]) else ray_o.detach().clone()
        self.ray_d = ray_d.detach().clone()
        self.dL = dL_dout.detach().clone() if backward else None
        args = (bg, means, scales, rots, opac, shs, int(sh_degree))

        def body():
            ctx.build(means, scales, rots, opac, scale_modifier, refit=refit)
            f = ctx.forward(self.ray_o, self.ray_d, *args, scale_modifier=scale_modifier, record_hits=backward, cap=cap)
            g = ctx.backward(self.ray_o, self.ray_d, *args, f["out"], self.dL, hits=f, scale_modifier=scale_modifier) if backward else None
            return f, g

        if refit:
            ctx.build(means, scales, rots, opac, scale_modifier, refit=False)
        for _ in range(2):
            body()
        torch.cuda.synchronize(ctx.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.fwd, self.grads = body()

    def run(self, ray_o=None, ray_d=None, dL_dout=None):
        """Copy this frame's rays / upstream gradients into the static buffers (device-to-device) and replay. Returns the
        forward dict and the gradient dict; their tensors are the graph's static outputs, overwritten by the next run()."""
        if ray_o is not None:
            self.ray_o.copy_(ray_o.reshape(-1)[:3] if self.ray_o.numel() == 3 else ray_o, non_blocking=True)
        if ray_d is not None:
            self.ray_d.copy_(ray_d, non_blocking=True)
        if dL_dout is not None and self.dL is not None:
            self.dL.copy_(dL_dout, non_blocking=True)
        self.graph.replay()
        return self.fwd, self.grads

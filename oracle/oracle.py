"""oracle/oracle.py — TEST INFRASTRUCTURE. ctypes front-ends for the parity checkers.

* ``Oracle``  — the plain-C restatement in ``lidar_rt_oracle.c`` (float or double build).
* ``Ref``     — ``oracle/_ref``: the reference's own forward.cu / backward.cu compiled unmodified
                as host code over the OptiX stand-in (``oracle/build_ref.sh``). Present only when
                it was built in a container that has ``/root/reference``; the prebuilt ``.so``
                files travel to the GPU box.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import this module. The product package never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_float, c_int, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

ORC_TRIANGLES = 1
ORC_BVH = 2
ORC_FIX_BG = 4
ORC_FLAT = 8          # analysis mode: one hit list from the original origin, depth = t (lidar_rt_oracle.c)


def build(quiet: bool = True) -> None:
    """Compile the C restatement (and oracle/_ref when the reference tree is present)."""
    subprocess.run(["make", "-C", HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a, ct):
    return None if a is None else a.ctypes.data_as(POINTER(ct))


class Oracle:
    """C restatement of the reference tracer (see lidar_rt_oracle.c header)."""

    def __init__(self, double: bool = False):
        name = "liblidar_rt_oracle_f64.so" if double else "liblidar_rt_oracle_f32.so"
        path = os.path.join(HERE, name)
        if not os.path.exists(path):
            build()
        self.lib = ctypes.CDLL(path)
        self.dt = np.float64 if double else np.float32
        self.ct = c_double if double else c_float
        assert self.lib.orc_real_size() == np.dtype(self.dt).itemsize
        self.threads = self.lib.orc_num_threads()

    def _a(self, x, shape=None):
        a = np.ascontiguousarray(np.asarray(x, dtype=self.dt))
        if shape is not None:
            a = a.reshape(shape)
        return a

    def _common(self, ray_o, ray_d, bg, means, scales, rots, opac, shs):
        ray_d = self._a(ray_d).reshape(-1, 3)
        R = ray_d.shape[0]
        ray_o = self._a(ray_o).reshape(-1, 3)
        stride = 3 if ray_o.shape[0] == R and R > 1 else (3 if ray_o.shape[0] == R else 0)
        if ray_o.shape[0] == 1 and R > 1:
            stride = 0
        means = self._a(means).reshape(-1, 3)
        P = means.shape[0]
        scales = self._a(scales).reshape(P, 2)
        rots = self._a(rots).reshape(P, 4)
        opac = self._a(opac).reshape(P)
        shs = self._a(shs).reshape(P, -1, 3)
        M = shs.shape[1]
        bg = self._a(bg).reshape(3)
        return R, ray_o, stride, ray_d, bg, P, means, scales, rots, opac, shs, M

    def forward(self, ray_o, ray_d, bg, means, scales, rots, opac, shs, sh_degree,
                flags: int = 0, cap: int = 64, scale_modifier: float = 1.0):
        R, ray_o, stride, ray_d, bg, P, means, scales, rots, opac, shs, M = self._common(
            ray_o, ray_d, bg, means, scales, rots, opac, shs)
        out = np.zeros((R, 9), self.dt)
        accum = np.zeros(P, self.dt)
        hit_list = np.full((R, cap), -1, np.int32)
        hit_cnt = np.zeros(R, np.int32)
        slot_cnt = np.zeros(R, np.int32)
        ct = self.ct
        self.lib.orc_forward(
            c_int(R), _ptr(ray_o, ct), c_int(stride), _ptr(ray_d, ct), _ptr(bg, ct), c_int(P),
            _ptr(means, ct), _ptr(scales, ct), _ptr(rots, ct), _ptr(opac, ct), _ptr(shs, ct),
            c_int(sh_degree), c_int(M), ct(scale_modifier), c_int(flags),
            _ptr(out, ct), _ptr(accum, ct), _ptr(hit_list, c_int), _ptr(hit_cnt, c_int), c_int(cap),
            _ptr(slot_cnt, c_int))
        return dict(out=out, accum_w=accum, hit_list=hit_list, hit_cnt=hit_cnt, slot_cnt=slot_cnt)

    def backward(self, ray_o, ray_d, bg, means, scales, rots, opac, shs, sh_degree, fwd_out, dL_dout,
                 flags: int = 0, scale_modifier: float = 1.0):
        R, ray_o, stride, ray_d, bg, P, means, scales, rots, opac, shs, M = self._common(
            ray_o, ray_d, bg, means, scales, rots, opac, shs)
        fwd_out = self._a(fwd_out).reshape(R, 9)
        dL = self._a(dL_dout).reshape(R, 9)
        g = dict(means=np.zeros((P, 3), self.dt), shs=np.zeros((P, M, 3), self.dt), opac=np.zeros(P, self.dt),
                 scales=np.zeros((P, 2), self.dt), rots=np.zeros((P, 4), self.dt))
        ct = self.ct
        self.lib.orc_backward(
            c_int(R), _ptr(ray_o, ct), c_int(stride), _ptr(ray_d, ct), _ptr(bg, ct), c_int(P),
            _ptr(means, ct), _ptr(scales, ct), _ptr(rots, ct), _ptr(opac, ct), _ptr(shs, ct),
            c_int(sh_degree), c_int(M), ct(scale_modifier), c_int(flags),
            _ptr(fwd_out, ct), _ptr(dL, ct),
            _ptr(g["means"], ct), _ptr(g["shs"], ct), _ptr(g["opac"], ct), _ptr(g["scales"], ct), _ptr(g["rots"], ct))
        return g

    def build_rectangles(self, means, scales, rots, opac):
        means = self._a(means).reshape(-1, 3)
        P = means.shape[0]
        verts = np.zeros((4 * P, 3), self.dt)
        ct = self.ct
        self.lib.orc_build_rectangles(c_int(P), _ptr(means, ct), _ptr(self._a(scales).reshape(P, 2), ct),
                                      _ptr(self._a(rots).reshape(P, 4), ct), _ptr(self._a(opac).reshape(P), ct),
                                      _ptr(verts, ct))
        return verts

    def sh_eval(self, deg, direction, sh):
        c = np.zeros(3, self.dt)
        basis = np.zeros(16, self.dt)
        ct = self.ct
        self.lib.orc_sh_eval(c_int(deg), _ptr(self._a(direction).reshape(3), ct), _ptr(self._a(sh).reshape(-1, 3), ct),
                             _ptr(c, ct), _ptr(basis, ct))
        return c, basis


def ref_available() -> bool:
    return (os.path.exists(os.path.join(HERE, "_ref", "libref_forward.so"))
            and os.path.exists(os.path.join(HERE, "_ref", "libref_backward.so")))


class Ref:
    """oracle/_ref: the reference's forward.cu / backward.cu, compiled as host code."""

    def __init__(self):
        if not ref_available():
            raise RuntimeError("oracle/_ref is not built (needs /root/reference; run oracle/build_ref.sh)")
        self.fwd = ctypes.CDLL(os.path.join(HERE, "_ref", "libref_forward.so"))
        self.bwd = ctypes.CDLL(os.path.join(HERE, "_ref", "libref_backward.so"))
        self._orc = Oracle(False)

    @staticmethod
    def _f(x, shape):
        return np.ascontiguousarray(np.asarray(x, np.float32)).reshape(shape)

    def _prep(self, ray_o, ray_d, bg, means, scales, rots, opac, shs):
        ray_d = np.ascontiguousarray(np.asarray(ray_d, np.float32))
        H, W = (ray_d.shape[0], ray_d.shape[1]) if ray_d.ndim == 3 else (1, ray_d.reshape(-1, 3).shape[0])
        ray_d = ray_d.reshape(H * W, 3)
        ray_o = np.ascontiguousarray(np.broadcast_to(np.asarray(ray_o, np.float32).reshape(-1, 3), (H * W, 3)))
        means = self._f(means, (-1, 3)); P = means.shape[0]
        scales = self._f(scales, (P, 2)); rots = self._f(rots, (P, 4)); opac = self._f(opac, (P,))
        shs = self._f(shs, (P, -1, 3)); M = shs.shape[1]
        bg = self._f(bg, (3,))
        verts = self._orc.build_rectangles(means, scales, rots, opac)      # build2DRectangle restated
        return H, W, ray_o, ray_d, bg, P, means, scales, rots, opac, shs, M, verts

    def forward(self, ray_o, ray_d, bg, means, scales, rots, opac, shs, sh_degree, scale_modifier=1.0):
        H, W, ray_o, ray_d, bg, P, means, scales, rots, opac, shs, M, verts = self._prep(
            ray_o, ray_d, bg, means, scales, rots, opac, shs)
        out = np.zeros((H * W, 9), np.float32)
        accum = np.zeros(P, np.float32)
        f = c_float
        self.fwd.ref_forward(c_int(H), c_int(W), c_int(P), c_int(sh_degree), c_int(M),
                             _ptr(ray_o, f), _ptr(ray_d, f), _ptr(verts, f), _ptr(bg, f), _ptr(means, f), _ptr(shs, f),
                             _ptr(opac, f), _ptr(scales, f), c_float(scale_modifier), _ptr(rots, f),
                             _ptr(out, f), _ptr(accum, f))
        return dict(out=out, accum_w=accum)

    def backward(self, ray_o, ray_d, bg, means, scales, rots, opac, shs, sh_degree, fwd_out, dL_dout,
                 scale_modifier=1.0):
        H, W, ray_o, ray_d, bg, P, means, scales, rots, opac, shs, M, verts = self._prep(
            ray_o, ray_d, bg, means, scales, rots, opac, shs)
        fwd_out = self._f(fwd_out, (H * W, 9)); dL = self._f(dL_dout, (H * W, 9))
        g = dict(means=np.zeros((P, 3), np.float32), shs=np.zeros((P, M, 3), np.float32), opac=np.zeros(P, np.float32),
                 scales=np.zeros((P, 2), np.float32), rots=np.zeros((P, 4), np.float32))
        f = c_float
        self.bwd.ref_backward(c_int(H), c_int(W), c_int(P), c_int(sh_degree), c_int(M),
                              _ptr(ray_o, f), _ptr(ray_d, f), _ptr(verts, f), _ptr(bg, f), _ptr(means, f), _ptr(shs, f),
                              _ptr(opac, f), _ptr(scales, f), c_float(scale_modifier), _ptr(rots, f),
                              _ptr(fwd_out, f), _ptr(dL, f),
                              _ptr(g["means"], f), _ptr(g["shs"], f), _ptr(g["opac"], f), _ptr(g["scales"], f), _ptr(g["rots"], f))
        return g


class ChamferOracle:
    """C restatement of the reference's Chamfer kernels (chamfer_oracle.c; lib/utils/chamfer3D/chamfer3D.cu:11-196)."""

    def __init__(self):
        path = os.path.join(HERE, "libchamfer_oracle.so")
        if not os.path.exists(path):
            build()
        self.lib = ctypes.CDLL(path)
        self.threads = self.lib.chm_num_threads()

    @staticmethod
    def _pts(x):
        a = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
        if a.ndim == 2:
            a = a[None]
        assert a.ndim == 3 and a.shape[2] == 3
        return a

    def forward(self, xyz1, xyz2):
        """-> dist1 (b,n), dist2 (b,m), idx1 (b,n) int32, idx2 (b,m) int32 — the tuple chamfer_3DFunction.forward returns."""
        a, c = self._pts(xyz1), self._pts(xyz2)
        b, n, m = a.shape[0], a.shape[1], c.shape[1]
        assert c.shape[0] == b
        d1 = np.empty((b, n), np.float32); d2 = np.empty((b, m), np.float32)
        i1 = np.empty((b, n), np.int32); i2 = np.empty((b, m), np.int32)
        I = ctypes.c_int32
        self.lib.chm_forward(c_int(b), c_int(n), _ptr(a, c_float), c_int(m), _ptr(c, c_float), _ptr(d1, c_float), _ptr(i1, I),
                             _ptr(d2, c_float), _ptr(i2, I))
        return d1, d2, i1, i2

    def backward(self, xyz1, xyz2, g1, g2, idx1, idx2):
        a, c = self._pts(xyz1), self._pts(xyz2)
        b, n, m = a.shape[0], a.shape[1], c.shape[1]
        g1 = np.ascontiguousarray(g1, np.float32).reshape(b, n); g2 = np.ascontiguousarray(g2, np.float32).reshape(b, m)
        idx1 = np.ascontiguousarray(idx1, np.int32).reshape(b, n); idx2 = np.ascontiguousarray(idx2, np.int32).reshape(b, m)
        ga = np.empty_like(a); gc = np.empty_like(c)
        I = ctypes.c_int32
        self.lib.chm_backward(c_int(b), c_int(n), _ptr(a, c_float), c_int(m), _ptr(c, c_float), _ptr(g1, c_float), _ptr(g2, c_float),
                              _ptr(idx1, I), _ptr(idx2, I), _ptr(ga, c_float), _ptr(gc, c_float))
        return ga, gc

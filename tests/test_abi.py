"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares; the host layer mirrors the reference's Python surface; nothing here needs a GPU."""
import ctypes
import inspect
import os
import re

import pytest
import torch

from conftest import ROOT

LIB = os.path.join(ROOT, "lidar-rt_b200", "csrc", "liblidar_rt_b200.so")
HEADER = os.path.join(ROOT, "include", "lidar_rt_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lrt_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build the library first: lidar-rt_b200/csrc/build.sh (or __graft_entry__.build())"
    lib = ctypes.CDLL(LIB)
    names = _declared()
    assert {"lrt_ctx_create", "lrt_ctx_destroy", "lrt_build", "lrt_refit", "lrt_forward", "lrt_backward",
            "lrt_last_error", "lrt_get_info", "lrt_get_permutation", "lrt_version"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lidar_rt_b200.h but not exported"
    lib.lrt_version.restype = ctypes.c_int
    assert lib.lrt_version() == 100


def test_context_creation_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from lidar_rt_b200 import native
    with pytest.raises(native.LrtError):
        native.Context()
    lib = native.load_library()
    h = ctypes.c_void_p()
    assert lib.lrt_ctx_create(0, ctypes.byref(h)) != 0
    assert b"CUDA" in lib.lrt_last_error(None) or b"device" in lib.lrt_last_error(None)


def test_python_surface_matches_reference_signatures():
    import diff_lidar_tracer as dlt
    assert dlt.TracingSettings._fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier",
                                           "viewmatrix", "projmatrix", "sh_degree", "campos", "prefiltered", "debug")
    t = dlt.Tracer()                                    # no arguments, no GPU needed to construct
    assert isinstance(t, torch.nn.Module) and hasattr(t, "training")
    fwd = list(inspect.signature(t.forward).parameters)
    assert fwd == ["ray_o", "ray_d", "mesh_normals", "means3D", "grads3D", "shs", "colors_precomp", "opacities", "scales",
                   "rotations", "cov3Ds_precomp", "tracer_settings"]
    assert list(inspect.signature(t.build_acceleration_structure).parameters) == ["vertices", "triangles", "rebuild"]
    z = torch.zeros(1, 3)
    with pytest.raises(Exception, match="excatly one of either SHs"):
        t(z, z, None, z, z, shs=None, colors_precomp=None, opacities=z, scales=z, rotations=z)
    with pytest.raises(Exception, match="exactly one of either scale/rotation"):
        t(z, z, None, z, z, shs=z, opacities=z, scales=None, rotations=None, cov3Ds_precomp=None)
    import lib.gaussian_renderer as gr
    assert list(inspect.signature(gr.raytracing).parameters) == ["frame", "gaussian_assets", "sensor", "background", "args",
                                                                 "scaling_modifier", "override_color", "decomp"]
    assert gr.render is gr.raytracing and isinstance(gr.tracer_2dgs, dlt.Tracer)

#!/usr/bin/env bash
# oracle/build_ref_optix.sh — TEST INFRASTRUCTURE (baseline B1 / parity pin, never the product path).
# Compiles the UNMODIFIED reference tracer (submodules/diff-lidar-tracer) for the GPU from the sources
# where they lie under /root/reference, without its CMake/setup.py build:
#   * forward.cu / backward.cu -> PTX with `nvcc -ptx` (what its CMake `CUDA_PTX_COMPILATION` does)
#   * common.cpp, optix_wrapper.cpp, trace_surfels.cpp, ext.cpp -> the pybind11/torch module `_C`
#     (what its setup.py CUDAExtension does), via torch.utils.cpp_extension.load
# Outputs go ONLY to oracle/_ref_optix/ (git-ignored; travels to the GPU box).  Running it needs the
# driver's libnvoptix.so.1, which only exists on the GPU box: see oracle/run_ref_optix.py.
# Does nothing (exit 0) when the reference tree is absent, e.g. on the GPU box.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${LIDAR_RT_REFERENCE:-/root/reference}/submodules/diff-lidar-tracer"
OUT="$HERE/_ref_optix"
if [ ! -f "$REF/optix_tracer/forward.cu" ]; then
    echo "build_ref_optix: reference tree not found at $REF — skipping"
    exit 0
fi
if [ -f "$OUT/_C.so" ] && [ -f "$OUT/forward.ptx" ] && [ -f "$OUT/backward.ptx" ] && [ -z "${LRT_FORCE_REF_OPTIX:-}" ]; then
    echo "build_ref_optix: up to date ($OUT)"
    exit 0
fi
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
INC="-I$REF/optix_tracer -I$REF/third_party/glm -I$REF/third_party/optix/include"
# no --use_fast_math, as in the reference's CMake build (Release: -O3)
for f in forward backward; do
    "$NVCC" -ptx -O3 -std=c++17 -arch=compute_100 -w $INC "$REF/optix_tracer/$f.cu" -o "$OUT/$f.ptx"
done
REF="$REF" OUT="$OUT" python - <<'EOF'
import os, shutil, glob
os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
from torch.utils.cpp_extension import load
ref, out = os.environ["REF"], os.environ["OUT"]
bd = os.path.join(out, "build"); os.makedirs(bd, exist_ok=True)
load(name="_C", sources=[os.path.join(ref, s) for s in
        ("optix_tracer/common.cpp", "optix_tracer/optix_wrapper.cpp", "trace_surfels.cpp", "ext.cpp")],
     extra_include_paths=[os.path.join(ref, "third_party/optix/include"), os.path.join(ref, "third_party/glm"), ref],
     extra_cflags=["-O2", "-w"], extra_ldflags=["-lcuda", "-L/usr/local/cuda/lib64/stubs", "-ldl"],
     with_cuda=True, build_directory=bd, is_python_module=False, verbose=False)
shutil.copy(os.path.join(bd, "_C.so"), os.path.join(out, "_C.so"))
shutil.rmtree(bd, ignore_errors=True)
EOF
echo "build_ref_optix: built $OUT/_C.so $OUT/forward.ptx $OUT/backward.ptx"

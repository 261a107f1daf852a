#!/usr/bin/env bash
# Builds the pybind11 module `_C` (ext_b200.cpp: the reference's diff_lidar_tracer._C surface over the C ABI), in-tree:
#   lidar-rt_b200/diff_lidar_tracer/_C.so   (host C++ only: links libtorch and liblidar_rt_b200.so through an $ORIGIN rpath)
# Reused when newer than its sources.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../diff_lidar_tracer/_C.so"
if [ -f "$OUT" ] && [ "$OUT" -nt "$HERE/ext_b200.cpp" ] && [ "$OUT" -nt "$HERE/../../include/lidar_rt_b200.h" ] && [ -z "${LRT_FORCE_EXT:-}" ]; then
    echo "build_ext: up to date ($OUT)"; exit 0
fi
[ -f "$HERE/liblidar_rt_b200.so" ] || bash "$HERE/build.sh"
HERE="$HERE" OUT="$OUT" python - <<'PY'
import os, shutil
os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
from torch.utils.cpp_extension import load
here, out = os.environ["HERE"], os.environ["OUT"]
bd = os.path.join(here, "build", "ext"); os.makedirs(bd, exist_ok=True)
try:
    load(name="_C", sources=[os.path.join(here, "ext_b200.cpp")], extra_cflags=["-O2"],
         # ($$ -> $ by ninja, quotes for the shell) found next to the package (diff_lidar_tracer/../csrc) and from the test copy under oracle/_ref_pkg/diff_lidar_tracer
         extra_ldflags=[f"-L{here}", "-l:liblidar_rt_b200.so", "-Wl,-rpath,'$$ORIGIN/../csrc:$$ORIGIN/../../../lidar-rt_b200/csrc'"],
         with_cuda=True, build_directory=bd, is_python_module=False, verbose=False)
except OSError:
    pass          # built; load() then dlopens it from the build directory, where the $ORIGIN rpath does not resolve
shutil.copy(os.path.join(bd, "_C.so"), out)
PY
echo "build_ext: built $OUT"

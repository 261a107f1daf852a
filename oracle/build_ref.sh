#!/usr/bin/env bash
# oracle/build_ref.sh — TEST INFRASTRUCTURE.
# Compiles the reference tracer's own device programs (forward.cu / backward.cu) as host C++,
# from the sources where they lie under /root/reference, against the OptiX stand-in in
# oracle/ref_shim/.  Outputs go ONLY to oracle/_ref/ (git-ignored; travels to the GPU box).
# Does nothing (exit 0) when the reference tree is absent, e.g. on the GPU box.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${LIDAR_RT_REFERENCE:-/root/reference}/submodules/diff-lidar-tracer"
OUT="$HERE/_ref"
if [ ! -f "$REF/optix_tracer/forward.cu" ]; then
    echo "build_ref: reference tree not found at $REF — skipping (prebuilt oracle/_ref is used if present)"
    exit 0
fi
mkdir -p "$OUT"
CUDA_INC="${CUDA_HOME:-/usr/local/cuda}/include"
# -ffp-contract=off: the reference PTX build may contract mul+add into FMA at nvcc's discretion;
#  a host build has no way to reproduce nvcc's choices, so we use the uncontracted IEEE form.
FLAGS="-O2 -std=c++17 -fPIC -shared -fopenmp -ffp-contract=off -w -Wl,-Bsymbolic -x c++"
INC="-I$HERE/ref_shim -I$REF/optix_tracer -I$REF/third_party/glm -I$CUDA_INC"
g++ $FLAGS $INC "$HERE/ref_shim/ref_forward.cpp"  -o "$OUT/libref_forward.so"
g++ $FLAGS $INC "$HERE/ref_shim/ref_backward.cpp" -o "$OUT/libref_backward.so"
echo "build_ref: built $OUT/libref_forward.so $OUT/libref_backward.so"

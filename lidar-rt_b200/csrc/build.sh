#!/usr/bin/env bash
# Builds liblidar_rt_b200.so (the C-ABI library, include/lidar_rt_b200.h) for sm_100a, in-tree.
#  -fmad=false : IEEE evaluation as written — the hit-order arithmetic contract (lrt_common.cuh)
#  -lineinfo   : ncu source correlation
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
HOSTCXX="/usr/bin/g++"; [ -x "$HOSTCXX" ] || HOSTCXX="g++"
OUT="${LRT_OUT:-$HERE/liblidar_rt_b200.so}"
"$NVCC" -ccbin "$HOSTCXX" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
    -Xcompiler -fPIC -shared ${LRT_NVCC_EXTRA:-} \
    "$HERE/lrt_api.cu" "$HERE/lrt_build.cu" "$HERE/lrt_forward.cu" "$HERE/lrt_backward.cu" "$HERE/lrt_prepare.cu" "$HERE/lrt_rays.cu" \
    -o "$OUT" -lcudart
echo "built $OUT"

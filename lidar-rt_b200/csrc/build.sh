#!/usr/bin/env bash
# Builds liblidar_rt_b200.so (the C-ABI library, include/lidar_rt_b200.h) for sm_100a, in-tree.
#  -fmad=false : IEEE evaluation as written — the hit-order arithmetic contract (lrt_common.cuh)
#  -lineinfo   : ncu source correlation
# One object per source, compiled in parallel; an object is reused when it is newer than its source and every header.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
HOSTCXX="/usr/bin/g++"; [ -x "$HOSTCXX" ] || HOSTCXX="g++"
OUT="${LRT_OUT:-$HERE/liblidar_rt_b200.so}"
OBJ="${LRT_OBJ_DIR:-$HERE/build}"
EXTRA="${LRT_NVCC_EXTRA:-}"
mkdir -p "$OBJ"
# a change of flags invalidates every object
STAMP="$OBJ/.flags"; FLAGS_NOW="$NVCC $HOSTCXX $EXTRA"
if [ ! -f "$STAMP" ] || [ "$(cat "$STAMP")" != "$FLAGS_NOW" ]; then rm -f "$OBJ"/*.o; echo "$FLAGS_NOW" > "$STAMP"; fi
SRCS=(lrt_api lrt_build lrt_forward lrt_backward lrt_prepare lrt_rays lrt_chamfer lrt_adam lrt_densify)
pids=()
for s in "${SRCS[@]}"; do
    o="$OBJ/$s.o"; fresh=1
    if [ ! -f "$o" ]; then fresh=0; else
        for dep in "$HERE/$s.cu" "$HERE"/*.cuh "$HERE/../../include/lidar_rt_b200.h"; do [ "$dep" -nt "$o" ] && fresh=0; done
    fi
    if [ "$fresh" = 0 ]; then
        "$NVCC" -ccbin "$HOSTCXX" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
            -Xcompiler -fPIC $EXTRA -c "$HERE/$s.cu" -o "$o" &
        pids+=($!)
    fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -ccbin "$HOSTCXX" -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC \
    $(for s in "${SRCS[@]}"; do echo "$OBJ/$s.o"; done) -o "$OUT" -lcudart
echo "built $OUT"

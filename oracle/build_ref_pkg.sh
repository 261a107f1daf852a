#!/usr/bin/env bash
# oracle/build_ref_pkg.sh — TEST INFRASTRUCTURE. Assembles, under oracle/_ref_pkg/ (git-ignored; travels to the GPU box like
# oracle/_ref), a package directory `diff_lidar_tracer/` that holds
#   __init__.py   the reference's Python wrapper, UNMODIFIED, copied at build time from where it lies under /root/reference
#                 (submodules/diff-lidar-tracer/diff_lidar_tracer/__init__.py)
#   _C.so         THIS repository's pybind11 module (lidar-rt_b200/csrc/ext_b200.cpp -> lidar-rt_b200/diff_lidar_tracer/_C.so)
# so that tests/test_ref_wrapper.py can run the reference's own Tracer / _Tracer code over the B200 library.
# Does nothing (exit 0) when the reference tree is absent, e.g. on the GPU box (the assembled directory travels).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${LIDAR_RT_REFERENCE:-/root/reference}/submodules/diff-lidar-tracer/diff_lidar_tracer/__init__.py"
EXT="$HERE/../lidar-rt_b200/diff_lidar_tracer/_C.so"
OUT="$HERE/_ref_pkg/diff_lidar_tracer"
if [ ! -f "$REF" ]; then echo "build_ref_pkg: reference tree not found — skipping"; exit 0; fi
if [ ! -f "$EXT" ]; then echo "build_ref_pkg: $EXT not built (lidar-rt_b200/csrc/build_ext.sh) — skipping"; exit 0; fi
mkdir -p "$OUT"
cp "$REF" "$OUT/__init__.py"
cp "$EXT" "$OUT/_C.so"
echo "build_ref_pkg: assembled $OUT"

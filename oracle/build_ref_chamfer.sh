#!/usr/bin/env bash
# oracle/build_ref_chamfer.sh — TEST INFRASTRUCTURE (parity pin / speed baseline of SURVEY 8f N2, never the product path).
# Compiles the UNMODIFIED reference Chamfer extension (lib/utils/chamfer3D: chamfer_cuda.cpp + chamfer3D.cu) for the GPU
# from the sources where they lie under /root/reference, the way its own dist_chamfer_3D.py:18-22 does
# (torch.utils.cpp_extension.load), cross-compiled for sm_100. Output goes ONLY to oracle/_ref_chamfer/ (git-ignored;
# travels to the GPU box). Running it needs a GPU: see oracle/run_ref_chamfer.py.
# Does nothing (exit 0) when the reference tree is absent, e.g. on the GPU box.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${LIDAR_RT_REFERENCE:-/root/reference}/lib/utils/chamfer3D"
OUT="$HERE/_ref_chamfer"
if [ ! -f "$REF/chamfer3D.cu" ]; then
    echo "build_ref_chamfer: reference tree not found at $REF — skipping"
    exit 0
fi
if [ -f "$OUT/chamfer_3D.so" ] && [ -z "${LRT_FORCE_REF_CHAMFER:-}" ]; then
    echo "build_ref_chamfer: up to date ($OUT)"
    exit 0
fi
mkdir -p "$OUT"
REF="$REF" OUT="$OUT" python - <<'PY'
import os, shutil
os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
from torch.utils.cpp_extension import load
ref, out = os.environ["REF"], os.environ["OUT"]
bd = os.path.join(out, "build"); os.makedirs(bd, exist_ok=True)
load(name="chamfer_3D", sources=[os.path.join(ref, "chamfer_cuda.cpp"), os.path.join(ref, "chamfer3D.cu")],
     extra_cflags=["-O2", "-w"], extra_cuda_cflags=["-w"], with_cuda=True, build_directory=bd, is_python_module=False, verbose=False)
shutil.copy(os.path.join(bd, "chamfer_3D.so"), os.path.join(out, "chamfer_3D.so"))
shutil.rmtree(bd, ignore_errors=True)
PY
echo "build_ref_chamfer: built $OUT/chamfer_3D.so"

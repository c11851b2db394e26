#!/usr/bin/env bash
# oracle/build_ref_gpu.sh -- TEST INFRASTRUCTURE ONLY.
# Compiles the UNMODIFIED reference GPU library (GJK/gpu/openGJK.cu) from where it lies under $REF (default
# /root/reference) together with oracle/ref_gpu_driver.cu into oracle/_ref_gpu/libogjk_refgpu_f32.so: the reference's
# own kernels on this B200 = the second baseline ("kernel to beat") and a second oracle (reference-GPU against
# reference-CPU).  Flags are the reference's (GJK/CMakeLists.txt:32: --fmad=false) plus -O3; the code is built for
# sm_100 as a plain recompile -- that is the point of this baseline.  fp32 only (USE_32BITS is hard-defined at
# GJK/common.h:44).  Nothing from the reference is copied into the repository; the product never links this.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref_gpu"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
if [ ! -f "$REF/GJK/gpu/openGJK.cu" ]; then
  echo "build_ref_gpu: $REF not present; keeping prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
"$NVCC" -std=c++17 -O3 --fmad=false -gencode arch=compute_100,code=sm_100 -shared -Xcompiler -fPIC -w \
    -cudart static -I"$REF" -I"$REF/GJK/gpu" "$HERE/ref_gpu_driver.cu" "$REF/GJK/gpu/openGJK.cu" \
    -o "$OUT/libogjk_refgpu_f32.so"
echo "build_ref_gpu: wrote $OUT/libogjk_refgpu_f32.so"

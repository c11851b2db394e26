#!/usr/bin/env bash
# oracle/build_ref_vis.sh -- TEST INFRASTRUCTURE ONLY.
# Builds oracle/_ref_gpu/libogjk_refvis_f32.so: the UNMODIFIED kernels of the reference's visualiser that sit either
# side of the hot path (visualization/integrate_final_gjk.cu: quat_rotate, quat_rotate_inv, transform_to_world_kernel,
# insert_objects_kernel, count_pairs_kernel, generate_pairs_kernel, collision_response_kernel, init_polytopes_kernel)
# behind the host drivers of oracle/ref_vis_driver.cu.  The file they live in needs OpenGL headers, so the eight function
# definitions are extracted BY NAME into a mktemp scratch file (from the line that opens the definition to the first
# line that is a lone closing brace), compiled from there and the scratch file is deleted; nothing of the reference is
# copied into the repository.  Flags: nvcc defaults (FMA contraction on) -- what the reference's visualisation target
# uses; only its GJK library is built with --fmad=false (GJK/CMakeLists.txt:32).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref_gpu"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
SRC="$REF/visualization/integrate_final_gjk.cu"
if [ ! -f "$SRC" ]; then
  echo "build_ref_vis: $REF not present; keeping prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
extract() {  # $1 = regex matching the first line of the definition
  awk -v pat="$1" '$0 ~ pat {on=1} on {print} on && /^}/ {on=0; print ""}' "$SRC"
}
{
  extract '^__device__ static inline float3 quat_rotate\('
  extract '^__device__ static inline float3 quat_rotate_inv\('
  for k in transform_to_world_kernel insert_objects_kernel count_pairs_kernel generate_pairs_kernel \
           collision_response_kernel init_polytopes_kernel; do
    extract "^__global__ void $k\\("
  done
} > "$TMP/refvis_kernels.cuh"
test "$(grep -c '^__global__ void' "$TMP/refvis_kernels.cuh")" = 6
test "$(grep -c '^__device__ static inline float3 quat_rotate' "$TMP/refvis_kernels.cuh")" = 2
"$NVCC" -std=c++17 -O3 -gencode arch=compute_100,code=sm_100 -shared -Xcompiler -fPIC -w -cudart static \
    -I"$REF" -I"$REF/GJK/gpu" -DREFVIS_KERNELS="\"$TMP/refvis_kernels.cuh\"" "$HERE/ref_vis_driver.cu" \
    -o "$OUT/libogjk_refvis_f32.so"
echo "build_ref_vis: wrote $OUT/libogjk_refvis_f32.so"

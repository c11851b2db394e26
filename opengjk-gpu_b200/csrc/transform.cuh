// transform.cuh -- local -> world vertex transform and descriptor upkeep of the polytope pool.
//
// the step that rewrites every vertex each frame in the reference's caller
// (visualization/integrate_final_gjk.cu: transform_to_world_kernel :304-332, quat_rotate :102-113,
// init_polytopes_kernel :691-704).  Contract kept, per vertex v of sub-mesh sm owned by body b:
//     lv = v * scale[b] (component-wise);  rv = quat_rotate(quat[b], lv);  world = rv + pos[b]
// with quat_rotate(q = (u, s), v) = 2*dot(u,v)*u + (2*s*s - 1)*v + 2*s*cross(u,v), every expression evaluated in the
// source's left-to-right association with separately rounded fp32 operations (the numpy oracle does the same, so
// the world pool -- and therefore every GJK/EPA result downstream -- is reproducible bit for bit).
// Mapping: the reference runs one thread per sub-mesh that walks its vertices serially (20 000 threads, stride-12-byte
// accesses diverging across the warp); here a CTA takes one sub-mesh and its threads take consecutive vertices, so
// loads and stores of a warp are contiguous.
#pragma once
#include <cuda_runtime.h>

#include "gjk_math.cuh"
#include "ogjk_types.h"

namespace ogjk {

OGJK_D V3<float> quat_rotate_rn(const float4& q, const V3<float>& v) {
  const V3<float> u = mk<float>(q.x, q.y, q.z);
  const float s = q.w;
  const float dot_uv = dot(u, v);
  const V3<float> c = cross(u, v);
  const float two_dot = mul_rn(2.0f, dot_uv);
  const float k = sub_rn(mul_rn(mul_rn(2.0f, s), s), 1.0f);
  const float two_s = mul_rn(2.0f, s);
  return mk<float>(add_rn(add_rn(mul_rn(two_dot, u.x), mul_rn(k, v.x)), mul_rn(two_s, c.x)),
                   add_rn(add_rn(mul_rn(two_dot, u.y), mul_rn(k, v.y)), mul_rn(two_s, c.y)),
                   add_rn(add_rn(mul_rn(two_dot, u.z), mul_rn(k, v.z)), mul_rn(two_s, c.z)));
}

// vert_offsets / vert_counts / sub_mesh_body may be null: then every sub-mesh has `uniform_count` vertices at offset
// sm * uniform_count and belongs to body sm.
__global__ void __launch_bounds__(64)
transform_to_world_kernel(const float4* __restrict__ positions, const float4* __restrict__ quats,
                          const float* __restrict__ scales /* [bodies][3] */, const float* __restrict__ verts_local,
                          float* __restrict__ verts_world, const int* __restrict__ vert_offsets,
                          const int* __restrict__ vert_counts, const int* __restrict__ sub_mesh_body, int uniform_count,
                          int num_submeshes) {
  const int sm = blockIdx.x;
  if (sm >= num_submeshes) return;
  const int body = sub_mesh_body ? sub_mesh_body[sm] : sm;
  const long long offset = vert_offsets ? vert_offsets[sm] : (long long)sm * uniform_count;
  const int count = vert_counts ? vert_counts[sm] : uniform_count;
  const float4 pos = positions[body];
  const float4 q = quats[body];
  const V3<float> sc = mk<float>(scales[3 * body], scales[3 * body + 1], scales[3 * body + 2]);
  for (int v = threadIdx.x; v < count; v += blockDim.x) {
    const float* src = verts_local + 3 * (offset + v);
    const V3<float> lv = mk<float>(mul_rn(src[0], sc.x), mul_rn(src[1], sc.y), mul_rn(src[2], sc.z));
    const V3<float> rv = quat_rotate_rn(q, lv);
    float* dst = verts_world + 3 * (offset + v);
    dst[0] = add_rn(rv.x, pos.x);
    dst[1] = add_rn(rv.y, pos.y);
    dst[2] = add_rn(rv.z, pos.z);
  }
}

// descriptor upkeep (reference init_polytopes_kernel :691-704)
template <typename T>
__global__ void __launch_bounds__(256)
init_polytopes_kernel(PolytopeT<T>* __restrict__ polytopes, T* __restrict__ verts_world,
                      const int* __restrict__ vert_offsets, const int* __restrict__ vert_counts, int uniform_count,
                      int num_submeshes) {
  const int sm = blockIdx.x * blockDim.x + threadIdx.x;
  if (sm >= num_submeshes) return;
  const long long offset = vert_offsets ? vert_offsets[sm] : (long long)sm * uniform_count;
  PolytopeT<T> p;
  p.numpoints = vert_counts ? vert_counts[sm] : uniform_count;
  p.s[0] = p.s[1] = p.s[2] = T(0);
  p.s_idx = 0;
  p.coord = verts_world + 3 * offset;
  polytopes[sm] = p;
}

// Do the first n descriptors still describe the dense uniform layout the library remembers for this array
// (numpoints == nv, coord == base + i * nv * 3)?  Descriptor arrays are plain device memory that the caller may edit
// or re-point after upload (the reference allows it); the fast kernels read the remembered layout, so every call
// re-checks it and falls back to the descriptor-reading kernels when it no longer holds.
template <typename T>
__global__ void __launch_bounds__(256)
validate_dense_kernel(const PolytopeT<T>* __restrict__ desc, int n, const T* base, int nv, int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const PolytopeT<T> p = desc[i];
  if (p.numpoints != nv || p.coord != base + (size_t)i * nv * 3) *flag = 1;
}

}  // namespace ogjk

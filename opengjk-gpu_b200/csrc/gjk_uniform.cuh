// gjk_uniform.cuh -- the fast GJK kernel for uniform batches (every polytope of a body array has the same vertex
// count, coordinates dense [n][V][3] as the reference's flattening produces them, openGJK.cu:2928-2943).
//
// B200 mapping (what differs from the general kernel in gjk_generic.cuh):
//   * each pair's two vertex sets are read from HBM exactly once, with 128-bit loads straight from the API's
//     xyz-interleaved layout, and stay in REGISTERS for all GJK iterations: lane l of the L lanes that share a pair
//     owns vertices [l*VPL, (l+1)*VPL) of both bodies (the reference re-reads every vertex from global memory on
//     every iteration with three scalar loads, openGJK.cu:1209-1219);
//   * the max-dot scan uses Blackwell's packed fp32x2 multiply (FMUL2): the interleaved layout pairs up as
//     (x,y)*(dx,dy), (z,x')*(dz,dx), (y',z')*(dy,dz), so 4 vertices cost 6 FMUL2 + 8 FADD.  The adds stay scalar on
//     purpose: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false, which would break
//     bit-parity with the reference's unfused arithmetic;
//   * the per-lane maximum is a FMNMX3 tree, the L-lane reduction a value-only xor-butterfly; the owner lane and
//     the lowest index are recovered with one ballot and an equality scan only when the support point actually
//     improves (SURVEY.md Appendix A.2 semantics: lowest index among maxima, strictly above the current support).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gjk_core.cuh"
#include "gjk_generic.cuh"
#include "ogjk_types.h"

namespace ogjk {

// ---- vector loads of 16 bytes -----------------------------------------------------------------------------------
OGJK_D void load16(const float* p, float* out) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  out[0] = t.x;
  out[1] = t.y;
  out[2] = t.z;
  out[3] = t.w;
}
OGJK_D void load16(const double* p, double* out) {
  const double2 t = __ldg(reinterpret_cast<const double2*>(p));
  out[0] = t.x;
  out[1] = t.y;
}
template <typename T>
struct Vec16 {
  static constexpr int kElems = 16 / sizeof(T);
};

// products p[e] = v[e] * d[e % 3] for e in [0, 3*VPL)
template <int N>
OGJK_D void products(const float (&v)[N], const V3<float>& d, float (&p)[N]) {
  static_assert(N % 6 == 0, "whole pairs of vertices");
  const float dd[3] = {d.x, d.y, d.z};
#pragma unroll
  for (int e = 0; e < N; e += 2) {
    unsigned long long a, b, r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(v[e]), "f"(v[e + 1]));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(dd[e % 3]), "f"(dd[(e + 1) % 3]));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(p[e]), "=f"(p[e + 1]) : "l"(r));
  }
}
template <int N>
OGJK_D void products(const double (&v)[N], const V3<double>& d, double (&p)[N]) {
  const double dd[3] = {d.x, d.y, d.z};
#pragma unroll
  for (int e = 0; e < N; ++e) p[e] = mul_rn(v[e], dd[e % 3]);
}

template <typename T, int VPL>
OGJK_D T lane_dots(const T (&v)[3 * VPL], const V3<T>& d, T (&dots)[VPL]) {
  T p[3 * VPL];
  products(v, d, p);
#pragma unroll
  for (int j = 0; j < VPL; ++j) dots[j] = add_rn(add_rn(p[3 * j], p[3 * j + 1]), p[3 * j + 2]);
  T m = dots[0];
#pragma unroll
  for (int j = 1; j < VPL; ++j) m = fmax_(m, dots[j]);
  return m;
}

// Support search over register-resident vertices.  `base` = this lane's first vertex index.
template <typename T, int L, int VPL>
OGJK_D void support_registers(const T (&v)[3 * VPL], const T* __restrict__ coord, const V3<T>& d, int lane,
                              unsigned gmask, int gshift, V3<T>& sup, int& sup_idx) {
  T dots[VPL];
  const T mine = lane_dots<T, VPL>(v, d, dots);
  T best = mine;
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) best = fmax_(best, ShflT<T>::xor_(gmask, best, o));
  if (best > dot(sup, d)) {
    // owner = lowest lane of the group whose local maximum equals the global one (lanes own ascending index ranges)
    const unsigned eq = __ballot_sync(gmask, mine == best) & gmask;
    const int owner = __ffs(eq) - 1;  // warp lane id
    int k = VPL - 1;
#pragma unroll
    for (int j = VPL - 2; j >= 0; --j)
      if (dots[j] == best) k = j;
    const int idx = __shfl_sync(gmask, lane * VPL + k, owner);
    (void)gshift;
    const T* p = coord + 3 * (size_t)idx;
    sup = mk<T>(__ldg(p), __ldg(p + 1), __ldg(p + 2));
    sup_idx = idx;
  }
}

// Loads this lane's VPL vertices of one body.  Vertex groups past the end of the polytope are filled with copies
// of vertex 0, which can tie with but never beat the real vertex 0 (lowest index wins ties).
template <typename T, int VPL>
OGJK_D void load_lane_vertices(const T* __restrict__ body, int nv, int lane, const V3<T>& p0, T (&v)[3 * VPL]) {
  constexpr int E = Vec16<T>::kElems;
  const int first = lane * VPL;
#pragma unroll
  for (int g = 0; g < VPL / 4; ++g) {
    const int vtx = first + 4 * g;
    if (vtx + 4 <= nv) {
      const T* src = body + 3 * (size_t)vtx;
#pragma unroll
      for (int q = 0; q < 12 / E; ++q) load16(src + q * E, &v[12 * g + q * E]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool in = vtx + j < nv;
        const T* src = body + 3 * (size_t)(in ? vtx + j : 0);
        v[12 * g + 3 * j + 0] = in ? __ldg(src) : p0.x;
        v[12 * g + 3 * j + 1] = in ? __ldg(src + 1) : p0.y;
        v[12 * g + 3 * j + 2] = in ? __ldg(src + 2) : p0.z;
      }
    }
  }
}

template <typename T, int L, int VPL>
__global__ void __launch_bounds__(256)
gjk_uniform_kernel(const T* __restrict__ coord1, const T* __restrict__ coord2, int nv1, int nv2,
                   SimplexT<T>* __restrict__ simplices, T* __restrict__ distances, int n,
                   const uint32_t* __restrict__ tabs) {
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pair = gtid / L;
  if (pair >= n) return;
  const int lane = (int)(threadIdx.x & (L - 1));
  const int wlane = threadIdx.x & 31;
  const int gshift = wlane & ~(L - 1);
  const unsigned gmask = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << gshift);
  const T* __restrict__ c1 = coord1 + (size_t)pair * nv1 * 3;
  const T* __restrict__ c2 = coord2 + (size_t)pair * nv2 * 3;
  const uint32_t* t3 = tabs;
  const uint32_t* t2 = tabs + 4096;

  const V3<T> p0 = mk<T>(__ldg(c1), __ldg(c1 + 1), __ldg(c1 + 2));
  const V3<T> q0 = mk<T>(__ldg(c2), __ldg(c2 + 1), __ldg(c2 + 2));
  T va[3 * VPL], vb[3 * VPL];
  load_lane_vertices<T, VPL>(c1, nv1, lane, p0, va);
  load_lane_vertices<T, VPL>(c2, nv2, lane, q0, vb);

  GjkState<T> g;
  gjk_init(g, p0, q0);
  bool stop;
  do {
    ++g.k;
    // lanes hold group-relative vertex ranges; __shfl_sync wants warp lane ids, hence wlane for the index source
    support_registers<T, L, VPL>(va, c1, vneg(g.v), lane, gmask, gshift, g.sup1, g.idx1);
    support_registers<T, L, VPL>(vb, c2, g.v, lane, gmask, gshift, g.sup2, g.idx2);
    stop = gjk_advance(g, t2, t3);
  } while (!stop);

  if (lane == 0) {
    GlobalFetch<T> fetch{c1, c2};
    V3<T> w1, w2;
    gjk_witnesses(fetch, g.S, w1, w2);
    store_result(simplices + pair, distances + pair, g, w1, w2);
  }
}

}  // namespace ogjk

// gjk_generic.cuh -- the general GJK kernel: any vertex count per polytope, caller-built AoS
// descriptors (gkPolytope with a device `coord` pointer), direct or indexed pairs.
//
// This is the kernel behind every entry point that receives opaque device descriptors
// (compute_minimum_distance_device / _indexed_device, reference GJK/gpu/openGJK.h:237-243, 419-425).
// Work decomposition: L lanes (a power of two, 2..32) cooperate on one pair.  The O(V) support scans
// are strided over the L lanes and reduced with xor-shuffles to (max value, lowest index); everything
// else (exit tests, table-driven sub-algorithm, witnesses) is evaluated redundantly by the L lanes on
// register-resident state, so there is no per-iteration broadcast traffic (the reference re-broadcasts
// the simplex with 24 shuffles per iteration, openGJK.cu:862-881).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gjk_core.cuh"
#include "ogjk_types.h"

namespace ogjk {

template <typename T>
struct ShflT;
template <>
struct ShflT<float> {
  static OGJK_D float xor_(unsigned m, float v, int o) { return __shfl_xor_sync(m, v, o); }
};
template <>
struct ShflT<double> {
  static OGJK_D double xor_(unsigned m, double v, int o) { return __shfl_xor_sync(m, v, o); }
};

template <typename T>
struct GlobalFetch {
  const T* c1;
  const T* c2;
  OGJK_D V3<T> operator()(int body, int i) const {
    const T* c = body ? c2 : c1;
    return mk<T>(__ldg(c + 3 * (size_t)i), __ldg(c + 3 * (size_t)i + 1), __ldg(c + 3 * (size_t)i + 2));
  }
};

// support search of one body over L cooperating lanes: global maximum of
// dot(vertex, d), lowest index among ties, accepted only if strictly above the current support's value.
template <typename T, int L>
OGJK_D void support_strided(const T* __restrict__ coord, int n, const V3<T>& d, int lane, unsigned gmask,
                            V3<T>& sup, int& sup_idx) {
  T best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = lane; i < n; i += L) {
    const T* p = coord + 3 * (size_t)i;
    const T val = dot(__ldg(p), __ldg(p + 1), __ldg(p + 2), d);
    if (val > best) {
      best = val;
      bi = i;
    }
  }
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) {
    const T ov = ShflT<T>::xor_(gmask, best, o);
    const int oi = __shfl_xor_sync(gmask, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  if (best > dot(sup, d)) {
    const T* p = coord + 3 * (size_t)bi;
    sup = mk<T>(__ldg(p), __ldg(p + 1), __ldg(p + 2));
    sup_idx = bi;
  }
}

template <typename T>
OGJK_D void store_slot(SimplexT<T>* out, int j, const SV<T>& s, bool live) {
  out->vrtx[j][0] = live ? s.p.x : T(0);
  out->vrtx[j][1] = live ? s.p.y : T(0);
  out->vrtx[j][2] = live ? s.p.z : T(0);
  out->vrtx_idx[j][0] = live ? s.i1 : 0;
  out->vrtx_idx[j][1] = live ? s.i2 : 0;
}

template <typename T>
OGJK_D void store_result(SimplexT<T>* out, T* dist, const GjkState<T>& g, const V3<T>& w1, const V3<T>& w2) {
  out->nvrtx = g.S.n;
  store_slot(out, 0, g.S.s0, g.S.n > 0);
  store_slot(out, 1, g.S.s1, g.S.n > 1);
  store_slot(out, 2, g.S.s2, g.S.n > 2);
  store_slot(out, 3, g.S.s3, g.S.n > 3);
  out->witnesses[0][0] = w1.x;
  out->witnesses[0][1] = w1.y;
  out->witnesses[0][2] = w1.z;
  out->witnesses[1][0] = w2.x;
  out->witnesses[1][1] = w2.y;
  out->witnesses[1][2] = w2.z;
  *dist = sqrt_rn(norm2(g.v));
}

// ---- where a pair's two polytopes come from ---------------------------------------------------
template <typename T>
struct BodyRef {
  const T* c;
  int n;
};
// two parallel descriptor arrays (compute_minimum_distance_kernel, reference openGJK.cu:1427-1449)
template <typename T>
struct DescSource {
  const PolytopeT<T>* bd1;
  const PolytopeT<T>* bd2;
  OGJK_D void get(long long i, BodyRef<T>& a, BodyRef<T>& b) const {
    a.c = bd1[i].coord;
    a.n = bd1[i].numpoints;
    b.c = bd2[i].coord;
    b.n = bd2[i].numpoints;
  }
};
// one pool + (idx1, idx2) records (compute_minimum_distance_indexed_kernel, openGJK.cu:1451-1476)
template <typename T>
struct IndexedSource {
  const PolytopeT<T>* pool;
  const CollisionPair* pairs;
  OGJK_D void get(long long i, BodyRef<T>& a, BodyRef<T>& b) const {
    const CollisionPair pr = pairs[i];
    a.c = pool[pr.idx1].coord;
    a.n = pool[pr.idx1].numpoints;
    b.c = pool[pr.idx2].coord;
    b.n = pool[pr.idx2].numpoints;
  }
};
// dense n x nverts x 3 arrays, no descriptors
template <typename T>
struct UniformSource {
  const T* c1;
  const T* c2;
  int nv1, nv2;
  OGJK_D void get(long long i, BodyRef<T>& a, BodyRef<T>& b) const {
    a.c = c1 + (size_t)i * nv1 * 3;
    a.n = nv1;
    b.c = c2 + (size_t)i * nv2 * 3;
    b.n = nv2;
  }
};

template <typename T, int L, typename Source>
__global__ void __launch_bounds__(256)
gjk_generic_kernel(const Source src, SimplexT<T>* __restrict__ simplices, T* __restrict__ distances, int n,
                   const uint32_t* __restrict__ tabs) {
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pair = gtid / L;
  if (pair >= n) return;
  const int lane = (int)(threadIdx.x & (L - 1));
  const unsigned gmask = (L == 32) ? 0xffffffffu : (((1u << L) - 1u) << ((threadIdx.x & 31) & ~(L - 1)));

  BodyRef<T> A, B;
  src.get(pair, A, B);
  const T* __restrict__ c1 = A.c;
  const T* __restrict__ c2 = B.c;
  const int n1 = A.n, n2 = B.n;
  const uint32_t* t3 = tabs;
  const uint32_t* t2 = tabs + 4096;

  GjkState<T> g;
  gjk_init(g, mk<T>(__ldg(c1), __ldg(c1 + 1), __ldg(c1 + 2)), mk<T>(__ldg(c2), __ldg(c2 + 1), __ldg(c2 + 2)));
  bool stop;
  do {
    ++g.k;
    support_strided<T, L>(c1, n1, vneg(g.v), lane, gmask, g.sup1, g.idx1);
    support_strided<T, L>(c2, n2, g.v, lane, gmask, g.sup2, g.idx2);
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

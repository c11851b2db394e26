// epa_kernel.cuh -- Expanding Polytope Algorithm, one warp per colliding pair, polytope in shared memory.
//
// Behavioural contract: SURVEY.md Appendix A.6, i.e. the reference's scalar EPA (GJK/cpu/EPA.c:362-863), whose
// tie-breaks (lowest vertex index in the support search, lowest face slot in the closest-face search, lowest
// free slot for new faces in horizon-edge order) are reproduced exactly so results are bit-identical to it.
// What is different is the machine mapping:
//   * working set per pair is 4.6 KB (fp32) instead of the reference's 9.8 KB (openGJK.cu:1486-1501): at most
//     4 + 64 vertices can exist (one per iteration, EPA.c:592-597), per-corner provenance is looked up through
//     the vertex instead of being stored per face, vertex ids are bytes, everything is SoA so lane-strided
//     access is bank-conflict free;
//   * face planes are computed once, by the lane that creates the face (they depend only on the face's
//     vertices; the reference recomputes every live face each iteration, EPA.c:599-604);
//   * closest face, support, duplicate test, visibility, horizon uniqueness, slot assignment and face
//     construction are all lane-parallel; ordering decisions use ballots + popcounts instead of serial scans.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gjk_core.cuh"
#include "gjk_generic.cuh"
#include "ogjk_types.h"

namespace ogjk {

constexpr int kEpaMaxFaces = 128;  // reference EPA.c:43
constexpr int kEpaMaxIters = 64;   // reference EPA.c:592
constexpr int kEpaMaxVerts = 4 + kEpaMaxIters;

template <typename T>
struct EpaConfig;
template <>
struct EpaConfig<float> {
  static constexpr int kWarpsPerBlock = 8;
};
template <>
struct EpaConfig<double> {
  static constexpr int kWarpsPerBlock = 4;
};

template <typename T>
struct EpaWork {
  using real = T;
  static constexpr int kVerts = kEpaMaxVerts, kFaces = kEpaMaxFaces, kEdges = kEpaMaxFaces * 3, kRanks = kEpaMaxFaces;
  static constexpr int kMaxBodyVerts = 0x7fffffff;
  static constexpr bool kSmall = false, kLean = false;
  T vx[kEpaMaxVerts], vy[kEpaMaxVerts], vz[kEpaMaxVerts];  // Minkowski-difference vertices
  int src1[kEpaMaxVerts], src2[kEpaMaxVerts];              // provenance: vertex index on body 1 / body 2
  T nx[kEpaMaxFaces], ny[kEpaMaxFaces], nz[kEpaMaxFaces];  // unit outward normals
  T fd[kEpaMaxFaces];                                      // plane distances (>= 0)
  uint32_t fv[kEpaMaxFaces];                               // v0 | v1<<8 | v2<<16 | live<<24
  uint16_t edge[kEpaMaxFaces * 3];                         // scratch: directed edges of the dying faces, a<<8|b
  uint8_t rank2slot[kEpaMaxFaces];                         // scratch: r-th lowest free face slot
};

template <typename T>
OGJK_D V3<T> load3(const T* __restrict__ c, int i) {
  const T* p = c + 3 * (size_t)i;
  return mk<T>(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}

// EPA.c:350-360
template <typename T>
OGJK_D V3<T> normal_from_witnesses(const V3<T>& w1, const V3<T>& w2) {
  const V3<T> d = vsub(w2, w1);
  const T len = sqrt_rn(norm2(d));
  if (len > Tol<T>::eps()) return mk<T>(div_rn(d.x, len), div_rn(d.y, len), div_rn(d.z, len));
  return mk<T>(T(1), T(0), T(0));
}

// Warp-wide (max value, lowest index) reduction.  fp64: xor-butterfly on (value, index).  fp32: two REDUX
// instructions -- the maximum of an order-preserving integer key (-0 is first folded onto +0, which compare equal as
// floats), then the minimum index among the lanes that hold it.  Only the index is returned.
OGJK_D unsigned order_key(float x) {
  const unsigned b = __float_as_uint(add_rn(x, 0.0f));
  if (!(x == x)) return 0u;  // NaN never wins a strict '>' comparison
  return b ^ ((b & 0x80000000u) ? 0xffffffffu : 0x80000000u);
}
OGJK_D void warp_argmax(double& best, int& bi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = ShflT<double>::xor_(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
}
OGJK_D void warp_argmax(float& best, int& bi) {
  const unsigned key = order_key(best);
  const unsigned top = __reduce_max_sync(0xffffffffu, key);
  bi = (int)__reduce_min_sync(0xffffffffu, key == top ? (unsigned)bi : 0x7fffffffu);
}

// This lane's share of a small body (<= 64 vertices: vertices lane and lane + 32), fetched once per pair so that the
// 10-60 support searches of an expansion do not go back to L1/L2 for them.
template <typename T>
struct LaneVerts {
  V3<T> p[2];
  bool cached;
};
template <typename T>
OGJK_D LaneVerts<T> cache_lane_verts(const BodyRef<T>& A, int lane) {
  LaneVerts<T> r;
  r.cached = A.n <= 64;
  r.p[0] = r.p[1] = mk<T>(T(0), T(0), T(0));
  if (r.cached) {
    if (lane < A.n) r.p[0] = load3(A.c, lane);
    if (lane + 32 < A.n) r.p[1] = load3(A.c, lane + 32);
  }
  return r;
}
// per-lane pass of the support search: max of sign * dot(vertex, d) over this lane's vertices, strict '>' from -1e10
// in ascending index order
template <typename T>
OGJK_D void lane_support(const BodyRef<T>& A, const LaneVerts<T>& L, const V3<T>& d, bool negate, int lane, T& best,
                         int& bi) {
  best = (T)-1e10f;
  bi = 0x7fffffff;
  if (L.cached) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = lane + 32 * k;
      T sv = dot(L.p[k].x, L.p[k].y, L.p[k].z, d);
      if (negate) sv = -sv;
      if (i < A.n && sv > best) {
        best = sv;
        bi = i;
      }
    }
  } else {
    for (int i = lane; i < A.n; i += 32) {
      const V3<T> p = load3(A.c, i);
      T sv = dot(p.x, p.y, p.z, d);
      if (negate) sv = -sv;
      if (sv > best) {
        best = sv;
        bi = i;
      }
    }
  }
}

// EPA.c:307-344: Minkowski support from scratch; strict '>' from -1e10 in ascending index order, i.e. the
// lowest index attaining the maximum.  Returns false if either body has no vertex above -1e10.
template <typename T>
OGJK_D bool epa_support(const BodyRef<T>& A, const BodyRef<T>& B, const LaneVerts<T>& LA, const LaneVerts<T>& LB,
                        const V3<T>& d, int lane, V3<T>& w, int& i1, int& i2) {
  T b1, b2;
  int k1, k2;
  lane_support(A, LA, d, false, lane, b1, k1);
  lane_support(B, LB, d, true, lane, b2, k2);
  warp_argmax(b1, k1);
  warp_argmax(b2, k2);
  if (k1 == 0x7fffffff || k2 == 0x7fffffff) return false;
  w = vsub(load3(A.c, k1), load3(B.c, k2));
  i1 = k1;
  i2 = k2;
  return true;
}

// EPA.c:238-304: barycentric coordinates of the point of triangle (v0,v1,v2) closest to the origin
template <typename T>
OGJK_D void origin_barycentric(const V3<T>& v0, const V3<T>& v1, const V3<T>& v2, T& a0, T& a1, T& a2) {
  const V3<T> e0 = vsub(v1, v0), e1 = vsub(v2, v0);
  const T d00 = dot(e0, e0), d01 = dot(e0, e1), d11 = dot(e1, e1);
  const T d20 = -dot(v0, e0), d21 = -dot(v0, e1);
  const T denom = sub_rn(mul_rn(d00, d11), mul_rn(d01, d01));
  if (fabs_(denom) < Tol<T>::eps()) {
    a0 = a1 = a2 = div_rn(T(1), T(3));
    return;
  }
  const T inv = div_rn(T(1), denom);
  const T u = mul_rn(sub_rn(mul_rn(d11, d20), mul_rn(d01, d21)), inv);
  const T v = mul_rn(sub_rn(mul_rn(d00, d21), mul_rn(d01, d20)), inv);
  const T w = sub_rn(sub_rn(T(1), u), v);
  if (w < T(0)) {
    const V3<T> e12 = vsub(v2, v1);
    T t = div_rn(-dot(v1, e12), dot(e12, e12));
    t = fmax_(T(0), fmin_(T(1), t));
    a0 = T(0);
    a1 = sub_rn(T(1), t);
    a2 = t;
  } else if (u < T(0)) {
    T t = div_rn(-dot(v0, e1), dot(e1, e1));
    t = fmax_(T(0), fmin_(T(1), t));
    a0 = sub_rn(T(1), t);
    a1 = T(0);
    a2 = t;
  } else if (v < T(0)) {
    T t = div_rn(-dot(v0, e0), dot(e0, e0));
    t = fmax_(T(0), fmin_(T(1), t));
    a0 = sub_rn(T(1), t);
    a1 = t;
    a2 = T(0);
  } else {
    a0 = w;
    a1 = u;
    a2 = v;
  }
}

template <typename WT>
OGJK_D V3<typename WT::real> work_vertex(const WT& W, int i) {
  return mk<typename WT::real>(W.vx[i], W.vy[i], W.vz[i]);
}

// Build face `f` = (a, b, c): orient it away from the centroid (EPA.c:203-232, 791-819), then compute its plane
// (EPA.c:92-129).  Returns the packed vertex word with the live bit set; a degenerate face is reported through
// `degenerate` and is retired by the caller once all slots of this iteration are assigned.
template <typename WT>
OGJK_D uint32_t make_face(WT& W, int f, int a, int b, int c, const V3<typename WT::real>& centroid, bool& degenerate) {
  using T = typename WT::real;
  const V3<T> va = work_vertex(W, a), vb = work_vertex(W, b), vc = work_vertex(W, c);
  const V3<T> e0 = vsub(vb, va), e1 = vsub(vc, va);
  // cross(e0, e1) = (p - q) per component; after a swap of corners 1 and 2 the reference recomputes cross(e1, e0) =
  // (q - p) from the same products (EPA.c:92-129 after :203-232 / :791-819).  That is the exact negative EXCEPT for a
  // component that cancels to zero: p - p = +0 either way, where a negation would give -0.
  const T px = mul_rn(e0.y, e1.z), qx = mul_rn(e0.z, e1.y);
  const T py = mul_rn(e0.z, e1.x), qy = mul_rn(e0.x, e1.z);
  const T pz = mul_rn(e0.x, e1.y), qz = mul_rn(e0.y, e1.x);
  V3<T> nrm = mk<T>(sub_rn(px, qx), sub_rn(py, qy), sub_rn(pz, qz));
  if (dot(nrm, vsub(centroid, va)) > T(0)) {  // swap corners 1 and 2
    const int t = b;
    b = c;
    c = t;
    nrm = mk<T>(sub_rn(qx, px), sub_rn(qy, py), sub_rn(qz, pz));
  }
  const T len2 = norm2(nrm);
  const T eps = Tol<T>::eps();
  T d;
  if (len2 > mul_rn(eps, eps)) {
    const T len = sqrt_rn(len2);
    div3_rn(nrm.x, nrm.y, nrm.z, len);  // three correctly rounded quotients, one reciprocal
    d = dot(nrm, va);
    if (d < T(0)) {
      nrm = vneg(nrm);
      d = -d;
    }
    degenerate = false;
  } else {
    d = (T)1e10;
    degenerate = true;
  }
  W.nx[f] = nrm.x;
  W.ny[f] = nrm.y;
  W.nz[f] = nrm.z;
  W.fd[f] = d;
  return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | (1u << 24);
}

// Closest live face among slots [0, 32*nwords): smallest distance >= 0, lowest slot on ties (EPA.c:606-617).
// -1 if none.
OGJK_D void warp_argmin_nonneg(double& best, int& bf) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double od = ShflT<double>::xor_(0xffffffffu, best, o);
    const int of = __shfl_xor_sync(0xffffffffu, bf, o);
    if (od < best || (od == best && of < bf)) {
      best = od;
      bf = of;
    }
  }
}
OGJK_D void warp_argmin_nonneg(float& best, int& bf) {  // best >= 0 (or -0): raw bits order like the values
  const unsigned key = __float_as_uint(add_rn(best, 0.0f));
  const unsigned low = __reduce_min_sync(0xffffffffu, key);
  bf = (int)__reduce_min_sync(0xffffffffu, key == low ? (unsigned)bf : 0x7fffffffu);
  best = __shfl_sync(0xffffffffu, best, bf & 31);  // the winner's own value (keeps the sign of a zero distance)
}
template <typename T>
OGJK_D int closest_face(const EpaWork<T>& W, int lane, int nwords, T& dist) {
  T best = (T)1e10f;
  int bf = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < kEpaMaxFaces / 32; ++j) {
    if (j < nwords) {
      const int f = lane + 32 * j;
      const bool live = (W.fv[f] >> 24) != 0;
      const T d = W.fd[f];
      if (live && d >= T(0) && d < best) {
        best = d;
        bf = f;
      }
    }
  }
  warp_argmin_nonneg(best, bf);
  dist = best;
  return bf == 0x7fffffff ? -1 : bf;
}

// EPA for one pair, executed by one full warp with its private shared-memory work area.
template <typename T, typename Source>
OGJK_D void epa_pair(const Source& src, long long pair, EpaWork<T>& W, int lane, SimplexT<T>* __restrict__ simplices,
                     T* __restrict__ distances, T* __restrict__ normals) {
  SimplexT<T>* sp = simplices + pair;
  T* nrm_out = normals + 3 * (size_t)pair;
  const T eps = Tol<T>::eps();

  // ---- 1. gate (EPA.c:369-373): separated pairs only get a normal from the GJK witnesses ----------------
  const T dist_in = distances[pair];
  if (dist_in > eps) {
    if (lane == 0) {
      const V3<T> w1 = mk<T>(sp->witnesses[0][0], sp->witnesses[0][1], sp->witnesses[0][2]);
      const V3<T> w2 = mk<T>(sp->witnesses[1][0], sp->witnesses[1][1], sp->witnesses[1][2]);
      const V3<T> nr = normal_from_witnesses(w1, w2);
      nrm_out[0] = nr.x;
      nrm_out[1] = nr.y;
      nrm_out[2] = nr.z;
    }
    return;
  }

  BodyRef<T> A, B;
  src.get(pair, A, B);
  const LaneVerts<T> LA = cache_lane_verts(A, lane), LB = cache_lane_verts(B, lane);

  // simplex -> first vertices of the polytope
  const int nv_in = sp->nvrtx;
  int nv = nv_in;
  if (lane < 4) {
    W.vx[lane] = sp->vrtx[lane][0];
    W.vy[lane] = sp->vrtx[lane][1];
    W.vz[lane] = sp->vrtx[lane][2];
    W.src1[lane] = sp->vrtx_idx[lane][0];
    W.src2[lane] = sp->vrtx_idx[lane][1];
  }
  __syncwarp();

  // "no progress" exit shared by the regrow steps (EPA.c:409-418, 474-483, 559-567, 571-582)
  auto touch_exit = [&](int i1, int i2) {
    if (lane == 0) {
      const V3<T> w1 = load3(A.c, i1), w2 = load3(B.c, i2);
      const V3<T> nr = normal_from_witnesses(w1, w2);
      distances[pair] = T(0);
      sp->witnesses[0][0] = w1.x; sp->witnesses[0][1] = w1.y; sp->witnesses[0][2] = w1.z;
      sp->witnesses[1][0] = w2.x; sp->witnesses[1][1] = w2.y; sp->witnesses[1][2] = w2.z;
      nrm_out[0] = nr.x; nrm_out[1] = nr.y; nrm_out[2] = nr.z;
      sp->nvrtx = nv;
      for (int j = nv_in < 0 ? 0 : nv_in; j < nv && j < 4; ++j) {
        sp->vrtx[j][0] = W.vx[j]; sp->vrtx[j][1] = W.vy[j]; sp->vrtx[j][2] = W.vz[j];
        sp->vrtx_idx[j][0] = W.src1[j]; sp->vrtx_idx[j][1] = W.src2[j];
      }
    }
  };
  // candidate accepted iff at squared distance >= eps^2 from every current vertex (EPA.c:389-397 etc.)
  auto is_new = [&](const V3<T>& p) {
    const T eps_sq = mul_rn(eps, eps);
    bool fresh = true;
    for (int q = 0; q < nv; ++q) {
      const T dx = sub_rn(p.x, W.vx[q]), dy = sub_rn(p.y, W.vy[q]), dz = sub_rn(p.z, W.vz[q]);
      if (add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz)) < eps_sq) fresh = false;
    }
    return fresh;
  };
  auto push = [&](const V3<T>& p, int i1, int i2) {
    __syncwarp();
    if (lane == 0) {
      W.vx[nv] = p.x; W.vy[nv] = p.y; W.vz[nv] = p.z;
      W.src1[nv] = i1; W.src2[nv] = i2;
    }
    ++nv;
    __syncwarp();
  };

  // ---- 2. regrow a degenerate simplex to a tetrahedron (EPA.c:375-583) ---------------------------------------
  if (nv != 4) {
    V3<T> p;
    int i1 = 0, i2 = 0;
    if (nv == 1) {
      const bool ok = epa_support(A, B, LA, LB, work_vertex(W, 0), lane, p, i1, i2);
      if (ok && is_new(p)) push(p, i1, i2);
      else { touch_exit(i1, i2); return; }
    }
    if (nv == 2) {
      const V3<T> edge = vsub(work_vertex(W, 1), work_vertex(W, 0));
      V3<T> axis = mk<T>(T(1), T(0), T(0));
      const T len = sqrt_rn(norm2(edge));
      if (len > eps && fabs_(edge.x) > mul_rn((T)0.9f, len)) axis = mk<T>(T(0), T(1), T(0));
      V3<T> dir = cross(edge, axis);
      if (norm2(dir) < eps) dir = cross(edge, mk<T>(T(0), T(0), T(1)));
      const bool ok = epa_support(A, B, LA, LB, dir, lane, p, i1, i2);
      if (ok && is_new(p)) push(p, i1, i2);
      else { touch_exit(i1, i2); return; }
    }
    if (nv == 3) {
      const V3<T> v0 = work_vertex(W, 0);
      V3<T> dir = cross(vsub(work_vertex(W, 1), v0), vsub(work_vertex(W, 2), v0));
      bool ok = epa_support(A, B, LA, LB, dir, lane, p, i1, i2);
      if (ok && is_new(p)) {
        push(p, i1, i2);
      } else {
        dir = vneg(dir);
        ok = epa_support(A, B, LA, LB, dir, lane, p, i1, i2);
        if (ok && is_new(p)) push(p, i1, i2);
        else { touch_exit(i1, i2); return; }
      }
    }
    if (nv != 4) {  // nvrtx outside 1..4 on input (EPA.c:571-582)
      const int best = nv > 0 ? (nv - 1 < 3 ? nv - 1 : 3) : 0;
      touch_exit(sp->vrtx_idx[best][0], sp->vrtx_idx[best][1]);
      return;
    }
  }

  // ---- 3. tetrahedron (EPA.c:144-235) ---------------------------------------------------------------------------
  V3<T> centroid = mk<T>(T(0), T(0), T(0));
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    centroid.x = add_rn(centroid.x, mul_rn(W.vx[q], (T)0.25f));
    centroid.y = add_rn(centroid.y, mul_rn(W.vy[q], (T)0.25f));
    centroid.z = add_rn(centroid.z, mul_rn(W.vz[q], (T)0.25f));
  }
#pragma unroll
  for (int j = 0; j < kEpaMaxFaces / 32; ++j) W.fv[lane + 32 * j] = 0u;
  __syncwarp();
  {
    bool degenerate = false;
    uint32_t word = 0;
    if (lane < 4) {
      // faces (0,1,2) (0,3,1) (0,2,3) (1,3,2)
      const int a = lane == 3 ? 1 : 0;
      const int b = lane == 0 ? 1 : (lane == 2 ? 2 : 3);
      const int c = lane == 0 ? 2 : (lane == 1 ? 1 : (lane == 2 ? 3 : 2));
      word = make_face(W, lane, a, b, c, centroid, degenerate);
      W.fv[lane] = degenerate ? (word & 0x00ffffffu) : word;
    }
  }
  __syncwarp();

  // ---- 4. expansion (EPA.c:596-826) -----------------------------------------------------------------------------
  // New faces always take the lowest free slots (EPA.c:761-775), so live faces stay packed in the low slots: `hi`
  // (one past the highest slot ever used) bounds every per-face pass to ceil(hi/32) words instead of four.
  const T tol = Tol<T>::eps_tot();
  int iter = 0;
  int hi = 4;
  bool reported = false;
  int report_face = -1;
  T report_d = T(0);
  while (iter < kEpaMaxIters) {
    ++iter;
    const int nwords = (hi + 31) >> 5;
    T cd;
    const int cf = closest_face(W, lane, nwords, cd);
    if (cf < 0) break;
    const V3<T> cn = mk<T>(W.nx[cf], W.ny[cf], W.nz[cf]);
    V3<T> w;
    int i1 = 0, i2 = 0;
    if (!epa_support(A, B, LA, LB, cn, lane, w, i1, i2)) break;
    const T gain = sub_rn(dot(cn, w), cd);
    bool stop = gain < tol;
    if (!stop) {  // duplicate of an existing polytope vertex? (EPA.c:654-665)
      const T eps_sq = mul_rn(eps, eps);
      bool dup = false;
      for (int q = lane; q < nv; q += 32) {
        const T dx = sub_rn(w.x, W.vx[q]), dy = sub_rn(w.y, W.vy[q]), dz = sub_rn(w.z, W.vz[q]);
        if (add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz)) < eps_sq) dup = true;
      }
      stop = __any_sync(0xffffffffu, dup);
    }
    if (stop) {
      reported = true;
      report_face = cf;
      report_d = cd;
      break;
    }

    // add the vertex, move the running centroid (EPA.c:686-699)
    const int newv = nv;
    if (lane == 0) {
      W.vx[newv] = w.x; W.vy[newv] = w.y; W.vz[newv] = w.z;
      W.src1[newv] = i1; W.src2[newv] = i2;
    }
    ++nv;
    const T inv_n = div_rn(T(1), (T)nv);
    centroid.x = add_rn(centroid.x, mul_rn(sub_rn(w.x, centroid.x), inv_n));
    centroid.y = add_rn(centroid.y, mul_rn(sub_rn(w.y, centroid.y), inv_n));
    centroid.z = add_rn(centroid.z, mul_rn(sub_rn(w.z, centroid.z), inv_n));

    // faces that see the new vertex die; their directed edges go to the scratch list in (slot, corner) order
    uint32_t vis[kEpaMaxFaces / 32], fword[kEpaMaxFaces / 32];
    int nvis = 0;
#pragma unroll
    for (int j = 0; j < kEpaMaxFaces / 32; ++j) {
      vis[j] = 0u;
      fword[j] = 0u;
      if (j < nwords) {
        const int f = lane + 32 * j;
        const uint32_t word = W.fv[f];
        fword[j] = word;
        bool sees = false;
        if (word >> 24) {
          const int a = word & 0xff;
          const V3<T> diff = vsub(w, work_vertex(W, a));
          sees = dot(mk<T>(W.nx[f], W.ny[f], W.nz[f]), diff) > eps;
        }
        vis[j] = __ballot_sync(0xffffffffu, sees);
      }
    }
#pragma unroll
    for (int j = 0; j < kEpaMaxFaces / 32; ++j) {
      if (j < nwords) {
        if ((vis[j] >> lane) & 1u) {
          const int rank = nvis + __popc(vis[j] & ((1u << lane) - 1u));
          const uint32_t word = fword[j];
          const uint32_t a = word & 0xff, b = (word >> 8) & 0xff, c = (word >> 16) & 0xff;
          W.edge[3 * rank + 0] = (uint16_t)((a << 8) | b);
          W.edge[3 * rank + 1] = (uint16_t)((b << 8) | c);
          W.edge[3 * rank + 2] = (uint16_t)((c << 8) | a);
          fword[j] = word & 0x00ffffffu;
          W.fv[lane + 32 * j] = fword[j];  // retire
        }
        nvis += __popc(vis[j]);
      }
    }
    const int nedge = 3 * nvis;

    // r-th lowest free slot, for r < nedge: slots at or above `hi` are all free, so looking at [0, hi + nedge)
    // always finds enough -- unless the 128 slots run out, in which case the remaining edges are dropped (EPA.c:775)
    int nfree = 0;
    {
      const int limit = hi + nedge < kEpaMaxFaces ? hi + nedge : kEpaMaxFaces;
#pragma unroll
      for (int j = 0; j < kEpaMaxFaces / 32; ++j) {
        if (32 * j < limit) {
          const int f = lane + 32 * j;
          const bool is_free = (fword[j] >> 24) == 0;  // words at or above nwords were never loaded: 0 = free
          const uint32_t fm = __ballot_sync(0xffffffffu, is_free);
          if (is_free) W.rank2slot[nfree + __popc(fm & ((1u << lane) - 1u))] = (uint8_t)f;
          nfree += __popc(fm);
        }
      }
    }
    __syncwarp();

    // horizon = edges that occur exactly once (EPA.c:745-759); each gets the next lowest free slot, in edge
    // order (EPA.c:761-775)
    int base_rank = 0;
    bool any_degenerate = false;
    for (int e0 = 0; e0 < nedge; e0 += 32) {
      const int e = e0 + lane;
      bool keep = false;
      uint32_t key = 0;
      if (nedge <= 32) {
        // one edge per lane: an edge is on the horizon iff no other lane holds the same undirected edge
        uint32_t canon = 0xffff0000u | (uint32_t)lane;  // idle lanes: unique dummies
        if (e < nedge) {
          key = W.edge[e];
          const uint32_t x = key >> 8, y = key & 0xff;
          canon = x < y ? key : ((y << 8) | x);
        }
        const unsigned same = __match_any_sync(0xffffffffu, canon);
        keep = e < nedge && __popc(same) == 1;
      } else if (e < nedge) {
        key = W.edge[e];
        const uint32_t rev = ((key & 0xff) << 8) | (key >> 8);
        keep = true;
        for (int x = 0; x < nedge; ++x) {
          const uint32_t other = W.edge[x];
          if (x != e && (other == key || other == rev)) keep = false;
        }
      }
      const uint32_t keepm = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int q = base_rank + __popc(keepm & ((1u << lane) - 1u));
        if (q < nfree) {
          const int slot = W.rank2slot[q];
          bool degenerate = false;
          const uint32_t word = make_face(W, slot, (int)(key >> 8), (int)(key & 0xff), newv, centroid, degenerate);
          W.fv[slot] = word;
          if (degenerate) any_degenerate = true;
        }
      }
      base_rank += __popc(keepm);
    }
    {
      const int used = base_rank < nfree ? base_rank : nfree;
      if (used > 0) {
        const int top = (int)W.rank2slot[used - 1] + 1;  // ranks ascend with slots
        hi = top > hi ? top : hi;
      }
    }
    __syncwarp();
    // degenerate new faces stay "live" while slots are being handed out and are retired at the next plane
    // recomputation in the reference (EPA.c:125-128, 599-604) -- i.e. now.
    if (__any_sync(0xffffffffu, any_degenerate)) {
#pragma unroll
      for (int j = 0; j < kEpaMaxFaces / 32; ++j) {
        const int f = lane + 32 * j;
        if ((W.fv[f] >> 24) && W.fd[f] == (T)1e10) W.fv[f] &= 0x00ffffffu;
      }
      __syncwarp();
    }
  }

  // ---- 5. iteration cap: report the currently closest face (EPA.c:828-863) ----------------------------------
  if (!reported && iter >= kEpaMaxIters) {
    T cd;
    const int cf = closest_face(W, lane, kEpaMaxFaces / 32, cd);
    if (cf >= 0) {
      reported = true;
      report_face = cf;
      report_d = cd;
    }
  }

  // ---- outputs ------------------------------------------------------------------------------------------------------
  if (lane == 0) {
    if (nv_in != 4) {  // the regrown simplex is part of the result (EPA.c modifies it in place)
      sp->nvrtx = 4;
      for (int j = nv_in < 0 ? 0 : nv_in; j < 4; ++j) {
        sp->vrtx[j][0] = W.vx[j]; sp->vrtx[j][1] = W.vy[j]; sp->vrtx[j][2] = W.vz[j];
        sp->vrtx_idx[j][0] = W.src1[j]; sp->vrtx_idx[j][1] = W.src2[j];
      }
    }
    if (reported) {  // EPA.c:636-651
      const uint32_t word = W.fv[report_face];
      const int a = word & 0xff, b = (word >> 8) & 0xff, c = (word >> 16) & 0xff;
      T a0, a1, a2;
      origin_barycentric(work_vertex(W, a), work_vertex(W, b), work_vertex(W, c), a0, a1, a2);
      const V3<T> pa = load3(A.c, W.src1[a]), pb = load3(A.c, W.src1[b]), pc = load3(A.c, W.src1[c]);
      const V3<T> qa = load3(B.c, W.src2[a]), qb = load3(B.c, W.src2[b]), qc = load3(B.c, W.src2[c]);
      sp->witnesses[0][0] = add_rn(add_rn(mul_rn(pa.x, a0), mul_rn(pb.x, a1)), mul_rn(pc.x, a2));
      sp->witnesses[0][1] = add_rn(add_rn(mul_rn(pa.y, a0), mul_rn(pb.y, a1)), mul_rn(pc.y, a2));
      sp->witnesses[0][2] = add_rn(add_rn(mul_rn(pa.z, a0), mul_rn(pb.z, a1)), mul_rn(pc.z, a2));
      sp->witnesses[1][0] = add_rn(add_rn(mul_rn(qa.x, a0), mul_rn(qb.x, a1)), mul_rn(qc.x, a2));
      sp->witnesses[1][1] = add_rn(add_rn(mul_rn(qa.y, a0), mul_rn(qb.y, a1)), mul_rn(qc.y, a2));
      sp->witnesses[1][2] = add_rn(add_rn(mul_rn(qa.z, a0), mul_rn(qb.z, a1)), mul_rn(qc.z, a2));
      nrm_out[0] = W.nx[report_face];
      nrm_out[1] = W.ny[report_face];
      nrm_out[2] = W.nz[report_face];
      distances[pair] = -report_d;
    }
  }
}

// Gate + compaction (EPA.c:369-373): one thread per pair.  Separated pairs (distance > eps) get their contact
// normal from the GJK witnesses right here; colliding pairs are appended to a device-side queue so that only they
// occupy warps in epa_queue_kernel.  (The reference spends a full warp and 9.8 KB of shared memory on every
// pair, colliding or not: openGJK.cu:2729-2755.)
template <typename T>
__global__ void __launch_bounds__(256)
epa_gate_kernel(const SimplexT<T>* __restrict__ simplices, const T* __restrict__ distances, T* __restrict__ normals,
                int n, int* __restrict__ queue, int* __restrict__ counters) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool collide = false;
  if (i < n) {
    if (distances[i] > Tol<T>::eps()) {
      const SimplexT<T>* sp = simplices + i;
      const V3<T> w1 = mk<T>(sp->witnesses[0][0], sp->witnesses[0][1], sp->witnesses[0][2]);
      const V3<T> w2 = mk<T>(sp->witnesses[1][0], sp->witnesses[1][1], sp->witnesses[1][2]);
      const V3<T> nr = normal_from_witnesses(w1, w2);
      T* o = normals + 3 * (size_t)i;
      o[0] = nr.x;
      o[1] = nr.y;
      o[2] = nr.z;
    } else {
      collide = true;
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, collide);
  if (m) {
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&counters[0], __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (collide) queue[base + __popc(m & ((1u << lane) - 1u))] = (int)i;
  }
}

// Persistent EPA kernel: warps pull colliding pairs from the queue through an atomic ticket, so pairs whose
// expansion needs 60 iterations do not hold up those that need 3.  fp32: four 8-warp CTAs per SM (64 registers, 24
// bytes of spills) instead of three at 80 registers -- config 3: 7.64 against 8.44 ms per Mi pairs; five CTAs at 48
// registers lose again (9.2 ms).
template <typename T, typename Source>
__global__ void __launch_bounds__(EpaConfig<T>::kWarpsPerBlock * 32, sizeof(T) == 4 ? 4 : 1)
epa_queue_kernel(const Source src, SimplexT<T>* __restrict__ simplices, T* __restrict__ distances,
                 T* __restrict__ normals, const int* __restrict__ queue, int* __restrict__ counters) {
  __shared__ EpaWork<T> work[EpaConfig<T>::kWarpsPerBlock];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int count = counters[0];
  for (;;) {
    int q = 0;
    if (lane == 0) q = atomicAdd(&counters[1], 1);
    q = __shfl_sync(0xffffffffu, q, 0);
    if (q >= count) break;
    epa_pair<T, Source>(src, (long long)queue[q], work[warp], lane, simplices, distances, normals);
    __syncwarp();
  }
}

// One warp per pair, no compaction (used when every pair is expected to collide, or for tiny batches).
template <typename T, typename Source>
__global__ void __launch_bounds__(EpaConfig<T>::kWarpsPerBlock * 32)
epa_kernel(const Source src, SimplexT<T>* __restrict__ simplices, T* __restrict__ distances,
           T* __restrict__ normals, int n) {
  __shared__ EpaWork<T> work[EpaConfig<T>::kWarpsPerBlock];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pair = (long long)blockIdx.x * EpaConfig<T>::kWarpsPerBlock + warp;
  if (pair >= n) return;
  epa_pair<T, Source>(src, pair, work[warp], lane, simplices, distances, normals);
}

}  // namespace ogjk

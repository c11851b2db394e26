// epa_group.cuh -- Expanding Polytope Algorithm with a sub-warp GROUP of G lanes per colliding pair.
//
// Same algorithm, same arithmetic and the same tie-breaks as epa_pair (epa_kernel.cuh; contract: SURVEY.md Appendix
// A.6 = the reference's scalar EPA, GJK/cpu/EPA.c:362-863), but a different machine mapping.  profiles/
// r1c_epa_queue_cfg3.txt shows the warp-per-pair kernel issue-bound (77 % of the issue slots) with a flat profile:
// an expansion works on ~20 live faces, a handful of dying faces, <= 18 horizon edges and 4-6 new faces, so most of
// its ~560 warp instructions per iteration run with a fraction of the 32 lanes doing anything.  Here G = 4 (or 8) lanes
// share a pair and a warp carries EIGHT (four) pairs through one instruction stream:
//   * the kernel is a persistent state machine -- a group that finishes its pair pulls the next one from the queue
//     while the other groups of the warp keep expanding -- so all groups of a warp stay in one loop; reporting a pair
//     and setting up the next one are batched over the groups that need it (see the loop head);
//   * the expansion step is executed by all 32 lanes in lock step: collectives name the whole warp (each group looks
//     at its own G bits / shuffles within its G-lane segment), loops run to the maximum trip count over the groups
//     with per-group predicates.  (Collectives that name only the group's lanes are legal but ptxas serialises them
//     over the distinct masks -- the first version of this kernel ran at 16 active lanes and was slower than
//     warp-per-pair.)  Set-up and reporting, once per pair, stay group-masked and divergent;
//   * reductions are log2(G)-level xor butterflies on (value, lowest index);
//   * face slot f belongs to lane f % G of the group; per-face passes run over the live slot range only;
//   * both bodies' vertices are cached in registers (KV per lane and body: bodies of up to G * KV vertices).
// WPC warps per CTA (independent of one another: no CTA-level synchronisation).
//
// Work areas.  With the full-size EpaWork (4.7 KB, room for the reference's 64 iterations / 128 faces) only ~11 warps
// fit an SM and the kernel is latency-bound (profiles/r1e_experiments.txt).  The measured distribution of EPA
// iterations is short-tailed (config 3: mean 12.8, 99.5 % <= 24; configs 2 and 5: 99.9 % <= 25), so the areas used by
// default are cut to that: EpaWorkSmall (28 vertices, 56 face slots, 48 dying-face edges: 1.7 KB in fp32), EpaWorkTiny
// and EpaWorkLean (26 / 25 vertices, 1.4 KB).  A pair that would exceed any capacity of its area is abandoned
// untouched (EPA writes its outputs only when it reports) and appended to an overflow queue, which the warp-per-pair
// kernel with the full-size work area then processes from scratch; the result is therefore the same as if every pair
// had had the full-size area.  The kernel is latency-bound, so what the areas buy is occupancy: G = 4 runs 16 warps
// per SM on the 1.7 KB area (128 registers) and 20 on the 1.4 KB ones (96 registers); the launcher (ogjk_lib.cu,
// launch_epa_queue) has the policy and the measurements.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "epa_kernel.cuh"

namespace ogjk {

template <typename T>
struct EpaWorkSmall {
  using real = T;
  static constexpr int kVerts = 28, kFaces = 56, kEdges = 48, kRanks = 48;
  static constexpr int kMaxBodyVerts = 65535;  // provenance is kept in 16 bits
  static constexpr bool kSmall = true, kLean = false;
  T vx[kVerts], vy[kVerts], vz[kVerts];
  uint16_t src1[kVerts], src2[kVerts];
  T nx[kFaces], ny[kFaces], nz[kFaces];
  T fd[kFaces];
  uint32_t fv[kFaces];
  uint16_t edge[kEdges];
  uint8_t rank2slot[kRanks];  // free slots below `hi`, lowest first
};

// The same with everything cut to what 99 % of the pairs need (config 3: 0.8 % of the pairs run more than 23 iterations,
// configs 2 and 5: 0.3 % / 0.1 %): 26 vertices (22 expansions), 48 face slots, 12 dying faces per expansion, 16 free
// slots below `hi`, 8-bit provenance -- 1424 bytes (fp32), so that 20 warps fit an SM (five 4-warp CTAs at 96
// registers) where the 1.7 KB area allows 16.  Measured (profiles/r2y3_ab_epa_svc.txt): with bodies of up to 16 vertices
// (KV = 4: 96 registers without spills) 3.86 -> 3.43 ms per Mi pairs; with 32-vertex bodies (KV = 8 wants 120 registers:
// at 96 ptxas spills and the kernel is 20 % SLOWER -- 6.11 against 5.08 ms on config 3 -- while the area by itself costs
// nothing, 5.13 ms at 128 registers), so only the KV = 4 instantiation uses it.
// Bank layout: the eight groups of a warp read the same field of their own area in one instruction, four consecutive
// words per group, so the areas start 4 * odd banks apart (356 words = 4 mod 32; EpaWorkSmall: 428 = 12 mod 32).
template <typename T>
struct EpaWorkTiny {
  using real = T;
  static constexpr int kVerts = 26, kFaces = 48, kEdges = 36, kRanks = 16;
  static constexpr int kMaxBodyVerts = 255;  // provenance is kept in 8 bits
  static constexpr bool kSmall = true, kLean = false;
  T vx[kVerts], vy[kVerts], vz[kVerts];
  T nx[kFaces], ny[kFaces], nz[kFaces];
  T fd[kFaces];
  uint32_t fv[kFaces];
  alignas(8) uint16_t edge[kEdges];
  uint8_t src1[kVerts], src2[kVerts];
  uint8_t rank2slot[kRanks];
  uint8_t pad[sizeof(T) == 4 ? 12 : 4];
};
static_assert(sizeof(EpaWorkTiny<float>) == 1424 && (sizeof(EpaWorkTiny<float>) / 4) % 32 == 4, "bank layout");
static_assert((sizeof(EpaWorkSmall<float>) / 4) % 8 == 4, "bank layout");

// The LEAN area: EpaWorkTiny (one vertex less) plus the per-pair state that the kernel otherwise carries in registers
// across the expansion loop -- running centroid, the two body pointers, the report record, the pair index, and `nv` / `hi`,
// which are parked here while the support search has all 48 cached coordinates live.  Together with the bodies' vertex
// counts folded into a per-lane count of cached vertices (bodies that are not fully cached go to the overflow queue) and
// the iteration counter dropped (it is nv - 4), the 32-vertex instantiation (KV = 8) comes down from 120 to 96 registers
// with one to three local-memory loads per expansion step: five 4-warp CTAs = 20 warps per SM instead of 16.
// (Spills are poison here: shared memory takes nearly all of L1, so a local load is an L2 round trip on the warp's
// critical path -- the first 96-register build, ~30 spill instructions per step, was 20 % slower than the 128-register
// one, profiles/r2y3_ab_epa_svc.txt.  And nothing between 128 and 96 registers helps: a scheduler owns 16 K registers,
// i.e. 4 warps at 97..128 registers per thread and 5 at 96 -- two 9-warp CTAs at 112 registers ran ONE CTA per SM, 7.81 ms,
// and a 19-warp CTA at 104 did not launch, profiles/r2y6_ab_epa_svc.txt.)
// Measured (profiles/r2y7_ab_epa_svc.txt): config 3 5.09 -> 4.88 ms per Mi pairs, config 5 8.89 -> 8.40 ms per 3.96 M pairs;
// the same area at 128 registers / 16 warps costs 2 % (5.20 / 9.12: the smaller area's overflow and the parked state).
template <typename T>
struct EpaWorkLean {
  using real = T;
  static constexpr int kVerts = 25, kFaces = 46, kEdges = 36, kRanks = 12;
  static constexpr int kMaxBodyVerts = 255;
  static constexpr bool kSmall = true, kLean = true;
  T vx[kVerts], vy[kVerts], vz[kVerts];
  T nx[kFaces], ny[kFaces], nz[kFaces];
  T fd[kFaces];
  uint32_t fv[kFaces];
  alignas(8) uint16_t edge[kEdges];
  const T* body1;  // coordinates of the pair's two bodies
  const T* body2;
  T cx, cy, cz;    // running centroid
  uint8_t src1[kVerts], src2[kVerts];
  uint8_t rank2slot[kRanks];
  int8_t report_face;  // -1: nothing to report (the reported distance is fd[report_face])
  int8_t nv_in;
  uint8_t nv, hi;  // parked here while the support search has all 48 cached coordinates live
  int pair;
  int pad[sizeof(T) == 4 ? 6 : 1];  // fp32: 1424 bytes = 356 words = 4 mod 32 -- the eight groups' areas start in disjoint bank
                                    // quads (1400 bytes: config 5 8.39 instead of 8.32 ms, profiles/r2y8_ab_epa_svc.txt)
};
static_assert(sizeof(EpaWorkLean<float>) == 1424, "five 4-warp CTAs per SM: (32 areas + 1 KB) x 5 <= 228 KB");

template <int G>
struct Grp {
  static_assert(G == 4 || G == 8 || G == 16 || G == 32, "group size");
  int lane;       // lane within the group
  unsigned mask;  // warp lane mask of the group
  int shift;      // first warp lane of the group
  // whole_warp = false: collectives name only this group's lanes (legal anywhere the group is convergent, but ptxas
  // serialises votes over the distinct masks).  whole_warp = true: collectives name all 32 lanes -- only for code that
  // the four groups of a warp execute together; every group still sees just its own lanes' bits / values.
  OGJK_D explicit Grp(int wlane, bool whole_warp = false) {
    lane = wlane & (G - 1);
    shift = wlane & ~(G - 1);
    mask = (G == 32 || whole_warp) ? 0xffffffffu : (((1u << G) - 1u) << shift);
  }
  OGJK_D unsigned ballot(bool p) const {  // bit i = group lane i
    return (__ballot_sync(mask, p) >> shift) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
  }
  OGJK_D bool any(bool p) const { return ballot(p) != 0u; }
  OGJK_D void sync() const { __syncwarp(mask); }
  OGJK_D unsigned below() const { return (1u << lane) - 1u; }
  OGJK_D int bcast(int v, int src) const { return __shfl_sync(mask, v, src, G); }
};

// (max value, lowest index) over the group: xor butterfly; every lane ends with the same pair
template <typename T, int G>
OGJK_D void grp_argmax(const Grp<G>& g, T& best, int& bi) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    const T ov = __shfl_xor_sync(g.mask, best, o, G);
    const int oi = __shfl_xor_sync(g.mask, bi, o, G);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
}
template <typename T, int G>
OGJK_D void grp_argmin(const Grp<G>& g, T& best, int& bi) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    const T ov = __shfl_xor_sync(g.mask, best, o, G);
    const int oi = __shfl_xor_sync(g.mask, bi, o, G);
    if (ov < best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
}

// this lane's share of a body: vertices lane, lane + G, ... (cached in registers when the body has <= 64 vertices)
template <typename T, int G, int KV>
struct GrpVerts {
  static constexpr int kPerLane = KV;
  V3<T> p[kPerLane];
  bool cached;
};
template <typename T, int G, int KV>
OGJK_D void cache_grp_verts(const BodyRef<T>& A, int glane, GrpVerts<T, G, KV>& r) {
  r.cached = A.n <= G * KV;
#pragma unroll
  for (int k = 0; k < KV; ++k) {
    const int i = glane + G * k;
    r.p[k] = (r.cached && i < A.n) ? load3(A.c, i) : mk<T>(T(0), T(0), T(0));
  }
}
template <typename T, int G, int KV>
OGJK_D void grp_lane_support(const BodyRef<T>& A, const GrpVerts<T, G, KV>& L, const V3<T>& d, bool negate, int glane,
                             T& best, int& bi) {
  best = (T)-1e10f;
  bi = 0x7fffffff;
  if (L.cached) {
#pragma unroll
    for (int k = 0; k < KV; ++k) {
      const int i = glane + G * k;
      T sv = dot(L.p[k].x, L.p[k].y, L.p[k].z, d);
      if (negate) sv = -sv;
      if (i < A.n && sv > best) {
        best = sv;
        bi = i;
      }
    }
  } else {
    for (int i = glane; i < A.n; i += G) {  // per-lane trip counts: no collective inside
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
// the same for a body that is known to be cached; `cnt` = how many of this lane's KV slots hold a vertex
template <typename T, int G, int KV>
OGJK_D void grp_lane_support_cached(const GrpVerts<T, G, KV>& L, int cnt, const V3<T>& d, bool negate, int glane, T& best,
                                    int& bi) {
  best = (T)-1e10f;
  bi = 0x7fffffff;
#pragma unroll
  for (int k = 0; k < KV; ++k) {
    T sv = dot(L.p[k].x, L.p[k].y, L.p[k].z, d);
    if (negate) sv = -sv;
    if (k < cnt && sv > best) {
      best = sv;
      bi = glane + G * k;
    }
  }
}
// EPA.c:307-344 (see epa_support)
template <typename T, int G, int KV>
OGJK_D bool grp_support(const Grp<G>& g, const BodyRef<T>& A, const BodyRef<T>& B, const GrpVerts<T, G, KV>& LA,
                        const GrpVerts<T, G, KV>& LB, const V3<T>& d, V3<T>& w, int& i1, int& i2) {
  T b1, b2;
  int k1, k2;
  grp_lane_support<T, G, KV>(A, LA, d, false, g.lane, b1, k1);
  grp_lane_support<T, G, KV>(B, LB, d, true, g.lane, b2, k2);
  grp_argmax<T, G>(g, b1, k1);
  grp_argmax<T, G>(g, b2, k2);
  if (k1 == 0x7fffffff || k2 == 0x7fffffff) return false;
  w = vsub(load3(A.c, k1), load3(B.c, k2));
  i1 = k1;
  i2 = k2;
  return true;
}

// closest live face among slots [0, G*nj): smallest distance >= 0, lowest slot on ties (EPA.c:606-617); -1 if none
// `trips` >= nj is the loop bound (equal to nj, or the maximum over the warp's groups when they run in lock step)
template <typename T, int G, typename WT>
OGJK_D int grp_closest_face(const Grp<G>& g, const WT& W, int nj, int trips, T& dist) {
  T best = (T)1e10f;
  int bf = 0x7fffffff;
  for (int j = 0; j < trips; ++j) {
    const int f = g.lane + G * j;
    if (j < nj && f < WT::kFaces) {
      const bool live = (W.fv[f] >> 24) != 0;
      const T d = W.fd[f];
      if (live && d >= T(0) && d < best) {
        best = d;
        bf = f;
      }
    }
  }
  grp_argmin<T, G>(g, best, bf);
  dist = best;
  return bf == 0x7fffffff ? -1 : bf;
}

// KV: vertices per lane of each body cached in registers (bodies of up to G * KV vertices; larger ones are read from
// global memory on every support search -- or, in the lean instantiation, handed to the overflow queue).  WT: work area
// type (EpaWork, EpaWorkSmall, EpaWorkTiny, EpaWorkLean; WT::kLean selects the register-lean variant of the loop).
// counters: [0] queued pairs, [1] ticket, [2] overflow count (small work areas only: pairs appended to `overflow`).
// svc_batch / svc_defer: batching of the per-pair service code, see the loop head.
template <typename T, int G, int KV, typename WT, typename Source>
OGJK_D void epa_group_body(const Source& src, SimplexT<T>* __restrict__ simplices, T* __restrict__ distances,
                           T* __restrict__ normals, const int* __restrict__ queue, int* __restrict__ counters,
                           int* __restrict__ overflow, int svc_batch, int svc_defer) {
  extern __shared__ __align__(16) unsigned char epa_smem[];
  WT* work = reinterpret_cast<WT*>(epa_smem);
  const int wlane = threadIdx.x & 31;
  const Grp<G> g(wlane);
  WT& W = work[threadIdx.x / G];
  const T eps = Tol<T>::eps();
  const T tol = Tol<T>::eps_tot();

  enum { kIdle = 0, kExpand = 1, kReport = 2, kExit = 3 };
  int phase = kIdle;
  // per-pair state, replicated on the lanes of the group
  constexpr bool kLean = WT::kLean;
  int pair = 0;  // (the queue holds ints)
  SimplexT<T>* sp = nullptr;  // (lean: recomputed from `pair` where needed)
  T* nrm_out = nullptr;
  BodyRef<T> A, B;  // (lean: only alive during set-up; the loop uses the cached vertices, the report W.body1/2)
  A.c = B.c = nullptr;
  A.n = B.n = 0;
  GrpVerts<T, G, KV> LA, LB;
  LA.cached = LB.cached = false;
  unsigned vcnt = 0;  // lean: number of cached vertices of this lane, body 1 | body 2 << 8
  int nv_in = 0, nv = 0, hi = 4;  // (the iteration count of an expanding pair is nv - 4: every step but the last adds a vertex)
  V3<T> centroid = mk<T>(T(0), T(0), T(0));
  bool reported = false;
  int report_face = -1;
  T report_d = T(0);

  // Reporting a finished pair (lane 0 of its group, ~300 instructions) and setting up the next one (~500) run while the
  // other seven groups of the warp wait.  They are therefore BATCHED: a group that has finished waits until `svc_batch`
  // groups need service, or `svc_defer` expansion steps have passed, or nothing is expanding -- the service code then
  // runs once for all of them.
  int deferred = 0;
  for (;;) {
    bool service;
    {
      const unsigned need = __ballot_sync(0xffffffffu, phase == kIdle || phase == kReport);
      const unsigned busy = __ballot_sync(0xffffffffu, phase == kExpand);
      service = need != 0u && (busy == 0u || __popc(need) >= svc_batch * G || deferred >= svc_defer);
      deferred = service ? 0 : (need != 0u ? deferred + 1 : 0);
    }
    // ================================ outputs ======================================================================
    if (service && phase == kReport) {
      if constexpr (kLean) {
        pair = W.pair;
        sp = simplices + pair;
        nrm_out = normals + 3 * (size_t)pair;
        if (g.lane == 0) {
          nv_in = W.nv_in;
          report_face = W.report_face;
          reported = report_face >= 0;
          report_d = reported ? W.fd[report_face] : T(0);
          A.c = W.body1;
          B.c = W.body2;
        }
      }
      if (g.lane == 0) {
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
      phase = kIdle;
    }
    // ================================ fetch + set up the next pair =================================================
    if (service && phase == kIdle) {
      int q = 0;
      if (g.lane == 0) q = atomicAdd(&counters[1], 1);
      q = g.bcast(q, 0);
      if (q >= *reinterpret_cast<const volatile int*>(&counters[0])) {  // (not kept in a register across the loop)
        phase = kExit;
      } else {
        pair = queue[q];
        sp = simplices + pair;
        nrm_out = normals + 3 * (size_t)pair;
        src.get(pair, A, B);
        cache_grp_verts<T, G, KV>(A, g.lane, LA);
        cache_grp_verts<T, G, KV>(B, g.lane, LB);
        g.sync();  // the previous pair's last reads of the work area are done
        bool oversize = WT::kSmall && (A.n > WT::kMaxBodyVerts || B.n > WT::kMaxBodyVerts);
        if constexpr (kLean) {  // the lean loop only knows the cached vertices
          oversize = oversize || A.n > G * KV || B.n > G * KV;
          const int c1 = (A.n - g.lane + G - 1) / G, c2 = (B.n - g.lane + G - 1) / G;  // vertices lane, lane + G, ... < n
          vcnt = (unsigned)(c1 < 0 ? 0 : (c1 > KV ? KV : c1)) | ((unsigned)(c2 < 0 ? 0 : (c2 > KV ? KV : c2)) << 8);
        }
        if (oversize && g.lane == 0) overflow[atomicAdd(&counters[2], 1)] = (int)pair;  // phase stays kIdle
        nv_in = sp->nvrtx;
        nv = nv_in;
        if (g.lane < 4) {
          W.vx[g.lane] = sp->vrtx[g.lane][0];
          W.vy[g.lane] = sp->vrtx[g.lane][1];
          W.vz[g.lane] = sp->vrtx[g.lane][2];
          W.src1[g.lane] = sp->vrtx_idx[g.lane][0];
          W.src2[g.lane] = sp->vrtx_idx[g.lane][1];
        }
        g.sync();

        // "no progress" exit shared by the regrow steps (EPA.c:409-418, 474-483, 559-567, 571-582)
        auto touch_exit = [&](int i1, int i2) {
          if (g.lane == 0) {
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
          for (int q2 = 0; q2 < nv; ++q2) {
            const T dx = sub_rn(p.x, W.vx[q2]), dy = sub_rn(p.y, W.vy[q2]), dz = sub_rn(p.z, W.vz[q2]);
            if (add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz)) < eps_sq) fresh = false;
          }
          return fresh;
        };
        auto push = [&](const V3<T>& p, int i1, int i2) {
          g.sync();
          if (g.lane == 0) {
            W.vx[nv] = p.x; W.vy[nv] = p.y; W.vz[nv] = p.z;
            W.src1[nv] = i1; W.src2[nv] = i2;
          }
          ++nv;
          g.sync();
        };

        // ---- regrow a degenerate simplex to a tetrahedron (EPA.c:375-583) -----------------------------------------
        bool ok_setup = !oversize;
        if (ok_setup && nv != 4) {
          V3<T> p;
          int i1 = 0, i2 = 0;
          if (ok_setup && nv == 1) {
            const bool ok = grp_support<T, G, KV>(g, A, B, LA, LB, work_vertex(W, 0), p, i1, i2);
            if (ok && is_new(p)) push(p, i1, i2);
            else { touch_exit(i1, i2); ok_setup = false; }
          }
          if (ok_setup && nv == 2) {
            const V3<T> edge = vsub(work_vertex(W, 1), work_vertex(W, 0));
            V3<T> axis = mk<T>(T(1), T(0), T(0));
            const T len = sqrt_rn(norm2(edge));
            if (len > eps && fabs_(edge.x) > mul_rn((T)0.9f, len)) axis = mk<T>(T(0), T(1), T(0));
            V3<T> dir = cross(edge, axis);
            if (norm2(dir) < eps) dir = cross(edge, mk<T>(T(0), T(0), T(1)));
            const bool ok = grp_support<T, G, KV>(g, A, B, LA, LB, dir, p, i1, i2);
            if (ok && is_new(p)) push(p, i1, i2);
            else { touch_exit(i1, i2); ok_setup = false; }
          }
          if (ok_setup && nv == 3) {
            const V3<T> v0 = work_vertex(W, 0);
            V3<T> dir = cross(vsub(work_vertex(W, 1), v0), vsub(work_vertex(W, 2), v0));
            bool ok = grp_support<T, G, KV>(g, A, B, LA, LB, dir, p, i1, i2);
            if (ok && is_new(p)) {
              push(p, i1, i2);
            } else {
              dir = vneg(dir);
              ok = grp_support<T, G, KV>(g, A, B, LA, LB, dir, p, i1, i2);
              if (ok && is_new(p)) push(p, i1, i2);
              else { touch_exit(i1, i2); ok_setup = false; }
            }
          }
          if (ok_setup && nv != 4) {  // nvrtx outside 1..4 on input (EPA.c:571-582)
            const int best = nv > 0 ? (nv - 1 < 3 ? nv - 1 : 3) : 0;
            touch_exit(sp->vrtx_idx[best][0], sp->vrtx_idx[best][1]);
            ok_setup = false;
          }
        }
        if (ok_setup) {
          // ---- tetrahedron (EPA.c:144-235) ---------------------------------------------------------------------------
          centroid = mk<T>(T(0), T(0), T(0));
#pragma unroll
          for (int q2 = 0; q2 < 4; ++q2) {
            centroid.x = add_rn(centroid.x, mul_rn(W.vx[q2], (T)0.25f));
            centroid.y = add_rn(centroid.y, mul_rn(W.vy[q2], (T)0.25f));
            centroid.z = add_rn(centroid.z, mul_rn(W.vz[q2], (T)0.25f));
          }
          for (int f = g.lane; f < WT::kFaces; f += G) W.fv[f] = 0u;
          g.sync();
          if (g.lane < 4) {
            // faces (0,1,2) (0,3,1) (0,2,3) (1,3,2)
            const int l = g.lane;
            const int a = l == 3 ? 1 : 0;
            const int b = l == 0 ? 1 : (l == 2 ? 2 : 3);
            const int c = l == 0 ? 2 : (l == 1 ? 1 : (l == 2 ? 3 : 2));
            bool degenerate = false;
            const uint32_t word = make_face(W, l, a, b, c, centroid, degenerate);
            W.fv[l] = degenerate ? (word & 0x00ffffffu) : word;
          }
          g.sync();
          hi = 4;
          reported = false;
          report_face = -1;
          report_d = T(0);
          if constexpr (kLean) {
            if (g.lane == 0) {
              W.cx = centroid.x; W.cy = centroid.y; W.cz = centroid.z;
              W.body1 = A.c;
              W.body2 = B.c;
              W.report_face = -1;
              W.nv_in = (int8_t)nv_in;
              W.nv = 4;
              W.hi = 4;
              W.pair = pair;
            }
            g.sync();
          }
          phase = kExpand;
        }
      }
    }
    if (__all_sync(0xffffffffu, phase == kExit)) break;

    // ================================ one expansion step (EPA.c:596-826) ==========================================
    // Executed by ALL lanes of the warp in lock step (collectives name the whole warp, loops run to the maximum trip
    // count over the four groups); `act` = this lane's group is expanding.  A group that stops mid-way clears `act`
    // and idles to the end of the block.
    if (__any_sync(0xffffffffu, phase == kExpand)) {
      const Grp<G> gw(wlane, true);
      bool act = phase == kExpand;
      if (WT::kVerts >= kEpaMaxVerts && __any_sync(0xffffffffu, act && nv - 4 >= kEpaMaxIters)) {  // iteration cap (EPA.c:828-863): rare;
        const bool cap = act && nv - 4 >= kEpaMaxIters;  // an area with fewer vertices hands the pair over long before
        T cd;
        constexpr int kAllTrips = (WT::kFaces + G - 1) / G;
        const int cf = grp_closest_face<T, G, WT>(gw, W, cap ? kAllTrips : 0, kAllTrips, cd);
        if (cap) {
          if (cf >= 0) {
            if constexpr (kLean) {
              if (g.lane == 0) {
                W.report_face = (int8_t)cf;
              }
            } else {
              reported = true;
              report_face = cf;
              report_d = cd;
            }
          }
          phase = kReport;
          act = false;
        }
      }
      if constexpr (kLean) hi = W.hi;
      int nj = act ? (hi + G - 1) / G : 0;
      const int njw = __reduce_max_sync(0xffffffffu, nj);
      T cd;
      int cf = grp_closest_face<T, G, WT>(gw, W, nj, njw, cd);
      if (act && cf < 0) {
        phase = kReport;
        act = false;
      }
      if (!act) cf = 0;
      const V3<T> cn = mk<T>(W.nx[cf], W.ny[cf], W.nz[cf]);
      V3<T> w = mk<T>(T(0), T(0), T(0));
      int i1 = 0, i2 = 0;
      {
        T b1 = (T)-1e10f, b2 = (T)-1e10f;
        int k1 = 0x7fffffff, k2 = 0x7fffffff;
        if (act) {  // no collective inside; idle groups may hold stale body descriptors
          if constexpr (kLean) {
            grp_lane_support_cached<T, G, KV>(LA, (int)(vcnt & 0xffu), cn, false, g.lane, b1, k1);
            grp_lane_support_cached<T, G, KV>(LB, (int)(vcnt >> 8), cn, true, g.lane, b2, k2);
          } else {
            grp_lane_support<T, G, KV>(A, LA, cn, false, g.lane, b1, k1);
            grp_lane_support<T, G, KV>(B, LB, cn, true, g.lane, b2, k2);
          }
        }
        grp_argmax<T, G>(gw, b1, k1);
        grp_argmax<T, G>(gw, b2, k2);
        if (act && (k1 == 0x7fffffff || k2 == 0x7fffffff)) {
          phase = kReport;
          act = false;
        }
        if (act) {
          if constexpr (kLean) w = vsub(load3(W.body1, k1), load3(W.body2, k2));
          else w = vsub(load3(A.c, k1), load3(B.c, k2));
          i1 = k1;
          i2 = k2;
        }
      }
      if constexpr (kLean) {  // nothing of the step's bookkeeping was kept in registers across the support search
        asm volatile("" ::: "memory");
        hi = W.hi;
        nv = W.nv;
        nj = act ? (hi + G - 1) / G : 0;
        cd = W.fd[cf];  // (cf is 0 for idle groups)
      }
      bool stop = sub_rn(dot(cn, w), cd) < tol;
      {  // duplicate of an existing polytope vertex? (EPA.c:654-665)
        const T eps_sq = mul_rn(eps, eps);
        const int nvw = __reduce_max_sync(0xffffffffu, act ? nv : 0);
        bool dup = false;
        for (int q2 = g.lane; q2 < nvw; q2 += G) {
          if (act && q2 < nv) {
            const T dx = sub_rn(w.x, W.vx[q2]), dy = sub_rn(w.y, W.vy[q2]), dz = sub_rn(w.z, W.vz[q2]);
            if (add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz)) < eps_sq) dup = true;
          }
        }
        if (gw.any(dup)) stop = true;
      }
      if (act && stop) {
        if constexpr (kLean) {
          if (g.lane == 0) {
            W.report_face = (int8_t)cf;
          }
        } else {
          reported = true;
          report_face = cf;
          report_d = cd;
        }
        phase = kReport;
        act = false;
      }

      // small work area: a pair that needs more than its capacities is handed to the full-size kernel untouched
      bool ovf = false;
      if (WT::kSmall && act && nv >= WT::kVerts) {
        ovf = true;
        act = false;
      }

      // add the vertex, move the running centroid (EPA.c:686-699)
      const int newv = nv;
      if (act) {
        if (g.lane == 0) {
          W.vx[newv] = w.x; W.vy[newv] = w.y; W.vz[newv] = w.z;
          W.src1[newv] = i1; W.src2[newv] = i2;
        }
        ++nv;
        const T inv_n = div_rn(T(1), (T)nv);
        if constexpr (kLean) {
          if (g.lane == 0) {  // read by the face pass, behind the __syncwarp() in front of the horizon loop
            W.nv = (uint8_t)nv;
            const T ox = W.cx, oy = W.cy, oz = W.cz;
            W.cx = add_rn(ox, mul_rn(sub_rn(w.x, ox), inv_n));
            W.cy = add_rn(oy, mul_rn(sub_rn(w.y, oy), inv_n));
            W.cz = add_rn(oz, mul_rn(sub_rn(w.z, oz), inv_n));
          }
        } else {
          centroid.x = add_rn(centroid.x, mul_rn(sub_rn(w.x, centroid.x), inv_n));
          centroid.y = add_rn(centroid.y, mul_rn(sub_rn(w.y, centroid.y), inv_n));
          centroid.z = add_rn(centroid.z, mul_rn(sub_rn(w.z, centroid.z), inv_n));
        }
      }

      // faces that see the new vertex die; their directed edges go to the scratch list in (slot, corner) order.  An edge
      // a -> b is stored as  min | max << 8 | (a > b) << 15 : the low 15 bits are the UNDIRECTED edge, which is what the
      // horizon test compares.  The same pass ranks the free slots below `hi` (free before, or dying now) for the new
      // faces; slots at or above `hi` are all free and need no table.
      int nvis = 0, nfree_lo = 0;
      for (int j = 0; j < njw; ++j) {
        const int f = g.lane + G * j;
        uint32_t word = 0;
        bool sees = false, in_range = false;
        if (act && j < nj) {
          word = W.fv[f];
          in_range = f < hi;
          if (word >> 24) {
            const int a = word & 0xff;
            const V3<T> diff = vsub(w, work_vertex(W, a));
            sees = dot(mk<T>(W.nx[f], W.ny[f], W.nz[f]), diff) > eps;
          }
        }
        const unsigned vm = gw.ballot(sees);
        const bool free_now = in_range && ((word >> 24) == 0 || sees);
        const unsigned fm = gw.ballot(free_now);
        if (sees) {
          const int rank = nvis + __popc(vm & g.below());
          const uint32_t a = word & 0xff, b = (word >> 8) & 0xff, c = (word >> 16) & 0xff;
          if (3 * rank + 2 < WT::kEdges) {
            W.edge[3 * rank + 0] = (uint16_t)(a < b ? (a | (b << 8)) : (b | (a << 8) | 0x8000u));
            W.edge[3 * rank + 1] = (uint16_t)(b < c ? (b | (c << 8)) : (c | (b << 8) | 0x8000u));
            W.edge[3 * rank + 2] = (uint16_t)(c < a ? (c | (a << 8)) : (a | (c << 8) | 0x8000u));
          }
          W.fv[f] = word & 0x00ffffffu;  // retire
        }
        if (free_now) {
          const int r = nfree_lo + __popc(fm & g.below());
          if (!WT::kSmall || r < WT::kRanks) W.rank2slot[r] = (uint8_t)f;
        }
        nvis += __popc(vm);
        nfree_lo += __popc(fm);
      }
      if (WT::kSmall && act && (3 * nvis > WT::kEdges || nfree_lo > WT::kRanks)) {  // more dying faces than the lists hold
        ovf = true;
        act = false;
        nvis = 0;
      }
      const int nedge = 3 * nvis;  // 0 for inactive groups
      if (act && g.lane < 3 && (nedge & 3) && nedge + g.lane < ((nedge + 3) & ~3))  // pad to a whole 64-bit word
        W.edge[nedge + g.lane] = 0xffffu;
      const int nedgew = __reduce_max_sync(0xffffffffu, nedge);
      // r-th lowest free slot for r < nfree: the table below `hi`, then hi, hi + 1, ... -- unless the slots run out, in
      // which case the remaining edges are dropped (EPA.c:775)
      const int limit = !act ? 0 : (hi + nedge < WT::kFaces ? hi + nedge : WT::kFaces);
      const int nfree = !act ? 0 : nfree_lo + (limit - hi);
      auto free_slot = [&](int r) { return r < nfree_lo ? (int)W.rank2slot[r] : hi + (r - nfree_lo); };
      __syncwarp();

      // horizon = edges that occur exactly once, in either direction (EPA.c:745-759); each gets the next lowest free
      // slot, in edge order (EPA.c:761-775).  Four list entries per 64-bit load; no collective inside the inner loop, so
      // its trip count is the group's own.
      // (Measured and dropped: letting the whole warp test one group's edge list after the other with match.any on the
      // undirected key instead of this all-pairs loop -- config 3 went from 6.29 to 7.27 ms per Mi pairs,
      // profiles/r2d_ab_epa.txt: eight MATCH + VOTE + SHFL rounds per expansion cost more than the loop they replace.)
      int base_rank = 0;
      bool any_degenerate = false;
      const uint2* edge4 = reinterpret_cast<const uint2*>(W.edge);
      for (int e0 = 0; e0 < nedgew; e0 += G) {
        const int e = e0 + g.lane;
        bool keep = e < nedge;
        uint32_t key = 0xffffu;
        if (keep) {
          key = W.edge[e];
          const uint32_t und2 = (key & 0x7fffu) * 0x10001u;
          // per 16-bit half: (entry & 0x7fff) ^ key is 0 on a match; adding 0x7fff sets bit 15 of every NON-zero half
          // (no carry leaves a half: 0x7fff + 0x7fff < 0x10000); the halves that stay clear are the matches
          uint32_t miss = 0;
          int halves = 0;
          for (int x = 0; x < nedge; x += 4) {
            const uint2 four = edge4[x >> 2];
            miss += (((four.x & 0x7fff7fffu) ^ und2) + 0x7fff7fffu) >> 15 & 0x00010001u;
            miss += (((four.y & 0x7fff7fffu) ^ und2) + 0x7fff7fffu) >> 15 & 0x00010001u;
            halves += 4;
          }
          const int cnt = halves - (int)((miss & 0xffffu) + (miss >> 16));
          keep = cnt == 1;  // only itself
        }
        const unsigned keepm = gw.ballot(keep);
        if (keep) {
          const int q2 = base_rank + __popc(keepm & g.below());
          if (q2 < nfree) {
            const int slot = free_slot(q2);
            W.fv[slot] = key;  // parked in its (free: live byte 0) slot until the face pass below
          }
        }
        base_rank += __popc(keepm);
      }
      {
        // The horizon edges are now parked one per new slot, in rank order: the faces are built in a second pass whose
        // trip count is the number of NEW faces (dying faces + 2) instead of the number of dying-face edges (3 per face)
        // -- measured: config 3 5.49 -> 5.41 ms per Mi pairs (profiles/r2x_ab_epa_svc.txt).
        const int used_now = base_rank < nfree ? base_rank : nfree;
        const int usedw = __reduce_max_sync(0xffffffffu, used_now);
        __syncwarp();
        for (int r0 = 0; r0 < usedw; r0 += G) {
          const int r = r0 + g.lane;
          if (r < used_now) {
            if constexpr (kLean) centroid = mk<T>(W.cx, W.cy, W.cz);
            const int slot = free_slot(r);
            const uint32_t key = W.fv[slot];
            const int lo = (int)(key & 0xffu), hi8 = (int)((key >> 8) & 0x7fu);
            const bool rev = (key & 0x8000u) != 0u;
            bool degenerate = false;
            const uint32_t word = make_face(W, slot, rev ? hi8 : lo, rev ? lo : hi8, newv, centroid, degenerate);
            W.fv[slot] = word;
            if (degenerate) any_degenerate = true;
          }
        }
      }
      if (WT::kSmall && act && base_rank > nfree) {  // ran out of the small area's face slots: the full-size area has more
        ovf = true;
        act = false;
      }
      if (WT::kSmall && gw.any(ovf)) {  // (warp-uniform branch: every group looks at its own lanes' bits)
        if (ovf) {
          if constexpr (kLean) pair = W.pair;
          if (g.lane == 0) overflow[atomicAdd(&counters[2], 1)] = (int)pair;
          phase = kIdle;
        }
      }
      if (act) {
        const int used = base_rank < nfree ? base_rank : nfree;
        if (used > 0) {
          const int top = free_slot(used - 1) + 1;  // ranks ascend with slots
          hi = top > hi ? top : hi;
          if constexpr (kLean) {
            if (g.lane == 0) W.hi = (uint8_t)hi;
          }
        }
      }
      __syncwarp();
      // degenerate new faces stay "live" while slots are being handed out and are retired at the next plane
      // recomputation in the reference (EPA.c:125-128, 599-604) -- i.e. now.
      if (__any_sync(0xffffffffu, any_degenerate)) {
        const bool mine = gw.any(any_degenerate);
        if (mine)
          for (int f = g.lane; f < WT::kFaces; f += G)
            if ((W.fv[f] >> 24) && W.fd[f] == (T)1e10) W.fv[f] &= 0x00ffffffu;
        __syncwarp();
      }
    }

  }
}

// The register budget is given either as resident warps per SM (MINB: launch bound of MINB / WPC CTAs of WPC warps) ...
template <typename T, int G, int KV, typename WT, int MINB, int WPC, typename Source>
__global__ void __launch_bounds__(32 * WPC, MINB / WPC)
epa_group_kernel(const Source src, SimplexT<T>* __restrict__ simplices, T* __restrict__ distances,
                 T* __restrict__ normals, const int* __restrict__ queue, int* __restrict__ counters,
                 int* __restrict__ overflow, int svc_batch, int svc_defer) {
  epa_group_body<T, G, KV, WT, Source>(src, simplices, distances, normals, queue, counters, overflow, svc_batch, svc_defer);
}
// ... or as registers per thread (with a minimum-CTA launch bound ptxas settles on 96 registers for anything between 17
// and 20 warps per SM; the lean instantiation needs exactly 104 to stay free of spills).
template <typename T, int G, int KV, typename WT, int REGS, int WPC, typename Source>
__global__ void __launch_bounds__(32 * WPC) __maxnreg__(REGS)
epa_group_kernel_regs(const Source src, SimplexT<T>* __restrict__ simplices, T* __restrict__ distances,
                      T* __restrict__ normals, const int* __restrict__ queue, int* __restrict__ counters,
                      int* __restrict__ overflow, int svc_batch, int svc_defer) {
  epa_group_body<T, G, KV, WT, Source>(src, simplices, distances, normals, queue, counters, overflow, svc_batch, svc_defer);
}

}  // namespace ogjk

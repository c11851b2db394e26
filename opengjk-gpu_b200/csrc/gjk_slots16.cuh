// gjk_slots16.cuh -- persistent slot GJK kernel whose slots hold the vertices as CENTRED, SCALED fp16 (half the bytes
// of gjk_slots.cuh's slots: 256 pairs of 64+64 vertices per SM instead of 128, two compute warps per scheduler instead
// of one) while every result stays bit-identical to the reference's fp32 arithmetic.
//
// The support search (reference openGJK.cu:1199-1251 / openGJK.c:615-639: maximum of the individually rounded fp32 dot
// products, lowest index on ties, the current support point kept unless strictly beaten) is split in two:
//   1. PRE-SCAN over the fp16 copy: packed HMUL2/HFMA2 dot products (1.5 instructions per vertex instead of 3.5),
//      per-block maxima with HMNMX2.  The copy is  ch = fl16(s * (c - c0))  with c0 the body's first vertex and
//      s = 16000 / max|c - c0|; the direction is scaled by a power of two t so that max|t d_j| is in [0.5, 1).
//   2. EXACT VERIFICATION of the candidates: every vertex whose approximate value is within `slack` of the approximate
//      maximum is re-evaluated from global memory (L2: the converter warps have just streamed it through) with the
//      reference's operations, in index order, under the reference's strict '>' rule.
// Guarantee (proved in DESIGN.md, checked by scripts/proto_fp16_slots.py and tests/test_fp16_prescan_bound.py): with
//      slack = sum_j |t d_j| * W_j * (1 + 2^-12) + 0.5,      W_j = 86.1 + 3.7e-7 * s * |c0_j|
// the lowest-index maximiser of the fp32 values and every vertex tying it are candidates: 86.1 covers twice the fp16
// error (five roundings of relative size 2^-11 on terms bounded by 16001 |t d_j|) and 3.7e-7 s |c0_j| twice the fp32
// rounding error of the reference's own evaluation (3 * 2^-24 relative to |d_j| max|c_j|).  So the support point -- and
// with it every later bit of the iteration -- is the reference's.  On the benchmark generator 1.19 vertices per scan
// are candidates (83 % of the scans: one).
//
// Roles (one CTA per SM, 8 + NC + 1 warps):
//   * compute warps: ONE THREAD PER PAIR as in gjk_slots.cuh (rotated loop, lane-uniform iteration gjk_substep_u);
//   * converter warps replace the TMA loader: they poll the slot flags, draw tickets, fetch the pair's two fp32 vertex
//     sets with 128-bit loads (lane l: four consecutive vertices), centre / scale / convert them and write the fp16
//     slot in blocks of eight vertices  x0..x7 | y0..y7 | z0..z7  (three 128-bit shared loads per block in the scan);
//     the bytes in flight that the TMA version kept in idle slots are held in the converters' registers;
//   * the finisher warp takes 9-word records (pair, simplex size, v, vertex indices), re-reads the <= 8 source vertices
//     from global memory/L2, rebuilds the simplex points (the same fp32 subtraction), and runs the witness stage, the
//     result stores and the fused EPA gate exactly as in gjk_slots.cuh.
#pragma once
#include <cuda_fp16.h>

#include "gjk_slots.cuh"

namespace ogjk {

constexpr int kS16Slots = 256;
constexpr int kS16ComputeWarps = 8;
constexpr int kS16RingRecords = 64;
constexpr int kS16RecWords = 9;       // pair | n | v.xyz | polytope index 1 | polytope index 2 | vertex indices (2 words)
constexpr int kS16ScratchWords = 25;  // finisher: 24 vertex coordinates per lane, odd stride
enum : unsigned { kS16Free = 0u, kS16Ready = 1u, kS16Exit = 2u };
constexpr uint32_t kS16PickBytes = 1024;  // converter batch lists: NC * P entries of 16 bytes
constexpr float kS16Scale = 16000.0f;
constexpr float kS16WConst = 86.1f;
constexpr float kS16WCentre = 3.7e-7f;

// header: c0 of body 1 (3 words) | W of body 1 (3) | c0 of body 2 (3) | W of body 2 (3); then the fp16 blocks
constexpr uint32_t kS16HeaderBytes = 48;
__host__ __device__ constexpr uint32_t s16_slot_bytes(int nb1, int nb2) {
  uint32_t units = kS16HeaderBytes / 16u + 3u * (uint32_t)(nb1 + nb2);
  if ((units & 1u) == 0) units += 1;  // odd 16-byte stride: conflict-free 128-bit loads, one slot per lane
  return units * 16u;
}
__host__ __device__ constexpr uint32_t s16_fixed_bytes() {
  // table | ctrl | pair_of | idx1_of | idx2_of | ring control (16 B) | ready flags | ring | finisher scratch | batch lists
  return kSlotTableBytes + (uint32_t)kS16Slots * 16u + 16u + (uint32_t)kS16RingRecords * 4u +
         (((uint32_t)kS16RingRecords * kS16RecWords * 4u + 15u) & ~15u) + ((32u * kS16ScratchWords * 4u + 15u) & ~15u) +
         kS16PickBytes;
}
__host__ __device__ constexpr uint32_t s16_smem_bytes(int nb1, int nb2) {
  return s16_fixed_bytes() + (uint32_t)kS16Slots * s16_slot_bytes(nb1, nb2);
}

OGJK_D __half2 as_h2(unsigned u) {
  __half2 h;
  *reinterpret_cast<unsigned*>(&h) = u;
  return h;
}
// approximate dot products of one block of eight vertices (pairs of vertices per half2); the SAME instruction
// sequence serves the block maxima and the per-vertex re-evaluation of a candidate block, so both see the same bits
OGJK_D void block_dots16(const uint4& X, const uint4& Y, const uint4& Z, __half2 dx, __half2 dy, __half2 dz, __half2& a0,
                         __half2& a1, __half2& a2, __half2& a3) {
  a0 = __hfma2(as_h2(Z.x), dz, __hfma2(as_h2(Y.x), dy, __hmul2_rn(as_h2(X.x), dx)));
  a1 = __hfma2(as_h2(Z.y), dz, __hfma2(as_h2(Y.y), dy, __hmul2_rn(as_h2(X.y), dx)));
  a2 = __hfma2(as_h2(Z.z), dz, __hfma2(as_h2(Y.z), dy, __hmul2_rn(as_h2(X.z), dx)));
  a3 = __hfma2(as_h2(Z.w), dz, __hfma2(as_h2(Y.w), dy, __hmul2_rn(as_h2(X.w), dx)));
}

// the search direction of one iteration, prepared once for both bodies (body 2 is scanned along +v, body 1 along -v)
struct Dir16 {
  __half2 x, y, z;   // fl16(t v), duplicated in both halves
  float qx, qy, qz;  // |t v_j|
  bool wide;         // |v| outside [2^-60, 2^123]: the scaling cannot be formed, every vertex is a candidate
};
OGJK_D Dir16 make_dir16(const V3<float>& v) {
  Dir16 D;
  const float m = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fabsf(v.z));
  const unsigned eb = (__float_as_uint(m) >> 23) & 0xffu;  // biased exponent: m in [2^(eb-127), 2^(eb-126))
  D.wide = eb < 67u || eb > 250u;
  const float t = __uint_as_float((253u - (D.wide ? 127u : eb)) << 23);  // 2^(126-eb): max|t v_j| in [0.5, 1)
  const float tx = v.x * t, ty = v.y * t, tz = v.z * t;                  // exact (power of two, no underflow of note)
  D.qx = fabsf(tx);
  D.qy = fabsf(ty);
  D.qz = fabsf(tz);
  D.x = __float2half2_rn(tx);
  D.y = __float2half2_rn(ty);
  D.z = __float2half2_rn(tz);
  return D;
}

// candidate bookkeeping of one body during the verification rounds
struct Cand16 {
  unsigned blocks;  // candidate blocks not yet opened
  unsigned verts;   // candidate vertices of the open block
  int blk;
};

// pre-scan of one body: block maxima, approximate maximum, threshold, candidate-block mask
template <int NB>
OGJK_D void prescan16(const uint4* __restrict__ blk, __half2 dx, __half2 dy, __half2 dz, float slack, bool wide,
                      Cand16& c, __half2& thr2) {
  __half2 bm[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const uint4 X = blk[3 * b], Y = blk[3 * b + 1], Z = blk[3 * b + 2];
    __half2 a0, a1, a2, a3;
    block_dots16(X, Y, Z, dx, dy, dz, a0, a1, a2, a3);
    bm[b] = __hmax2(__hmax2(a0, a1), __hmax2(a2, a3));
  }
  __half2 all = bm[0];
#pragma unroll
  for (int b = 1; b < NB; ++b) all = __hmax2(all, bm[b]);
  const float M = fmaxf(__low2float(all), __high2float(all));
  thr2 = __half2half2(__float2half_rd(M - slack));  // rounded DOWN: never above the real threshold
  unsigned mask = 0;
#pragma unroll
  for (int b = 0; b < NB; ++b) mask |= (__hge2_mask(bm[b], thr2) != 0u ? 1u : 0u) << b;
  c.blocks = wide ? ((1u << NB) - 1u) : mask;
  c.verts = 0;
  c.blk = 0;
}
// next candidate vertex of a body in index order, or -1
OGJK_D int next_candidate16(Cand16& c, const uint4* __restrict__ blk, __half2 dx, __half2 dy, __half2 dz, __half2 thr2,
                            bool wide) {
  if (c.verts == 0u && c.blocks != 0u) {
    c.blk = __ffs((int)c.blocks) - 1;
    c.blocks &= c.blocks - 1u;
    const uint4 X = blk[3 * c.blk], Y = blk[3 * c.blk + 1], Z = blk[3 * c.blk + 2];
    __half2 a0, a1, a2, a3;
    block_dots16(X, Y, Z, dx, dy, dz, a0, a1, a2, a3);
    const unsigned g0 = __hge2_mask(a0, thr2), g1 = __hge2_mask(a1, thr2), g2 = __hge2_mask(a2, thr2),
                   g3 = __hge2_mask(a3, thr2);
    const unsigned vm = (g0 & 1u) | ((g0 >> 15) & 2u) | ((g1 & 1u) << 2) | ((g1 >> 13) & 8u) | ((g2 & 1u) << 4) |
                        ((g2 >> 11) & 32u) | ((g3 & 1u) << 6) | ((g3 >> 9) & 128u);
    c.verts = wide ? 0xffu : vm;
  }
  if (c.verts == 0u) return -1;
  const int k = __ffs((int)c.verts) - 1;
  c.verts &= c.verts - 1u;
  return 8 * c.blk + k;
}

OGJK_D uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// NB1 / NB2: blocks of eight vertices per body (vertex counts 8 NB1 and 8 NB2, 2 (NB1 + NB2) <= 32 converter lanes).
// IDX: pairs are gkCollisionPair records into one pool.  NC converter warps (a divisor of 8), P pairs per converter batch.
template <int NB1, int NB2, bool IDX, int NC, int P>
__global__ void __launch_bounds__((kS16ComputeWarps + NC + 1) * 32)
gjk_slots16_kernel(const float* __restrict__ coord1, const float* __restrict__ coord2, SimplexT<float>* __restrict__ simplices,
                   float* __restrict__ distances, unsigned n, const uint16_t* __restrict__ utab_g,
                   unsigned* __restrict__ ticket, float* __restrict__ normals, int* __restrict__ epa_queue,
                   int* __restrict__ epa_count, const CollisionPair* __restrict__ pairs) {
  typedef float T;
  constexpr int CW = kS16ComputeWarps;
  constexpr int kThreads = (CW + NC + 1) * 32;
  constexpr int NV1 = 8 * NB1, NV2 = 8 * NB2;
  constexpr int G1 = 2 * NB1, G2 = 2 * NB2;  // converter lanes per body (four vertices each)
  static_assert(G1 + G2 <= 32, "a pair must fit the 32 lanes of a converter warp");
  static_assert(kS16Slots % (32 * NC) == 0, "every converter lane polls a whole number of slots");
  constexpr uint32_t sbytes = s16_slot_bytes(NB1, NB2);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* sp = smem_raw;
  uint16_t* utab = reinterpret_cast<uint16_t*>(sp);
  sp += kSlotTableBytes;
  unsigned* ctrl = reinterpret_cast<unsigned*>(sp);
  sp += kS16Slots * 4;
  unsigned* pair_of = reinterpret_cast<unsigned*>(sp);
  sp += kS16Slots * 4;
  int* idx1_of = reinterpret_cast<int*>(sp);
  sp += kS16Slots * 4;
  int* idx2_of = reinterpret_cast<int*>(sp);
  sp += kS16Slots * 4;
  unsigned* ring_ctl = reinterpret_cast<unsigned*>(sp);  // [0] tail (reserved), [1] head (consumed), [2] exited warps
  sp += 16;
  unsigned* ready = reinterpret_cast<unsigned*>(sp);
  sp += kS16RingRecords * 4;
  unsigned* ring = reinterpret_cast<unsigned*>(sp);
  sp += (kS16RingRecords * kS16RecWords * 4 + 15) & ~15;
  float* scratch = reinterpret_cast<float*>(sp);
  sp += (32 * kS16ScratchWords * 4 + 15) & ~15;
  uint4* picks = reinterpret_cast<uint4*>(sp);
  static_assert(NC * P * 16 <= (int)kS16PickBytes, "batch lists do not fit");
  unsigned char* slots = smem_raw + s16_fixed_bytes();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < kUnifiedSize / 2; i += kThreads)
    reinterpret_cast<uint32_t*>(utab)[i] = __ldg(reinterpret_cast<const uint32_t*>(utab_g) + i);
  for (int i = tid; i < kS16Slots; i += kThreads) {
    ctrl[i] = kS16Free;
    pair_of[i] = 0;
  }
  for (int i = tid; i < kS16RingRecords; i += kThreads) ready[i] = 0;
  if (tid < 4) ring_ctl[tid] = 0;
  __syncthreads();

  if (warp < CW) {
    // ================================================ compute ================================================
    const int cslot = tid;
    const unsigned char* sbase = slots + (size_t)cslot * sbytes;
    const float* hdr = reinterpret_cast<const float*>(sbase);
    const uint4* blk1 = reinterpret_cast<const uint4*>(sbase + kS16HeaderBytes);
    const uint4* blk2 = blk1 + 3 * NB1;
    enum { kWait = 0, kRun = 1, kExit = 2 };
    int state = kWait;
    unsigned pair = 0;
    int pi1 = 0, pi2 = 0;
    const float* g1 = coord1;  // this pair's fp32 vertices in global memory
    const float* g2 = coord2;
    GjkState<T> g;
    bool need_sub = false;  // rotated loop, see gjk_slots_ws_kernel
    for (;;) {
      bool fin_sub = false;
      if (state == kRun && need_sub) fin_sub = gjk_substep_u(g, utab);
      need_sub = false;
      if (state == kWait) {
        const unsigned c = ld_vol(&ctrl[cslot]);
        if (c == kS16Ready) {
          __threadfence_block();  // acquire: the converter's slot, header and pair_of stores
          pair = ld_vol(&pair_of[cslot]);
          pi1 = pi2 = (int)pair;
          if (IDX) {
            pi1 = *reinterpret_cast<volatile int*>(&idx1_of[cslot]);
            pi2 = *reinterpret_cast<volatile int*>(&idx2_of[cslot]);
          }
          g1 = coord1 + (size_t)pi1 * (NV1 * 3);
          g2 = coord2 + (size_t)pi2 * (NV2 * 3);
          gjk_init(g, mk<T>(hdr[0], hdr[1], hdr[2]), mk<T>(hdr[6], hdr[7], hdr[8]));
          state = kRun;
        } else if (c == kS16Exit) {
          state = kExit;
        }
      }
      if (__all_sync(0xffffffffu, state == kExit)) break;
      if (!__any_sync(0xffffffffu, state == kRun)) __nanosleep(32);  // start-up / drain
      bool finished = fin_sub;
      if (state == kRun && !fin_sub) {
        ++g.k;
        // ---- pre-scan of both bodies over the fp16 slot ----
        const Dir16 D = make_dir16(g.v);
        const __half2 nx = __hneg2(D.x), ny = __hneg2(D.y), nz = __hneg2(D.z);
        const float s1 = fmaf(D.qx, hdr[3], fmaf(D.qy, hdr[4], D.qz * hdr[5])) * 1.000244140625f + 0.5f;
        const float s2 = fmaf(D.qx, hdr[9], fmaf(D.qy, hdr[10], D.qz * hdr[11])) * 1.000244140625f + 0.5f;
        Cand16 c1, c2;
        __half2 thr1, thr2;
        prescan16<NB1>(blk1, nx, ny, nz, s1, D.wide, c1, thr1);
        prescan16<NB2>(blk2, D.x, D.y, D.z, s2, D.wide, c2, thr2);
        // ---- exact verification of the candidates, the reference's scan restricted to them ----
        const V3<T> nvv = vneg(g.v);
        T best1 = dot(g.sup1, nvv), best2 = dot(g.sup2, g.v);
        while ((c1.blocks | c1.verts | c2.blocks | c2.verts) != 0u) {
          const int i1 = next_candidate16(c1, blk1, nx, ny, nz, thr1, D.wide);
          const int i2 = next_candidate16(c2, blk2, D.x, D.y, D.z, thr2, D.wide);
          V3<T> p = g.sup1, q = g.sup2;
          if (i1 >= 0) {
            const float* a = g1 + 3 * i1;
            p = mk<T>(__ldg(a), __ldg(a + 1), __ldg(a + 2));
          }
          if (i2 >= 0) {
            const float* a = g2 + 3 * i2;
            q = mk<T>(__ldg(a), __ldg(a + 1), __ldg(a + 2));
          }
          if (i1 >= 0) {
            const T dd = dot(p, nvv);
            if (dd > best1) {
              best1 = dd;
              g.sup1 = p;
              g.idx1 = i1;
            }
          }
          if (i2 >= 0) {
            const T dd = dot(q, g.v);
            if (dd > best2) {
              best2 = dd;
              g.sup2 = q;
              g.idx2 = i2;
            }
          }
        }
        finished = gjk_converged_u(g);
        need_sub = !finished;
      }
      // ---- retire: 9-word record for the finisher, slot back to the converters ----
      {
        const unsigned fin = __ballot_sync(0xffffffffu, finished);
        if (fin) {
          const unsigned cnt = __popc(fin);
          unsigned base = 0;
          if (lane == 0) base = atomicAdd(&ring_ctl[0], cnt);
          base = __shfl_sync(0xffffffffu, base, 0);
          while ((int)(base + cnt - ld_vol(&ring_ctl[1])) > kS16RingRecords) __nanosleep(64);  // ring full
          const unsigned idx = base + __popc(fin & ((1u << lane) - 1u));
          if (finished) {
            unsigned* rec = ring + (size_t)(idx % kS16RingRecords) * kS16RecWords;
            rec[0] = pair;
            rec[1] = (unsigned)g.S.n;
            rec[2] = __float_as_uint(g.v.x);
            rec[3] = __float_as_uint(g.v.y);
            rec[4] = __float_as_uint(g.v.z);
            rec[5] = (unsigned)pi1;
            rec[6] = (unsigned)pi2;
            rec[7] = (unsigned)g.S.s0.i1 | ((unsigned)g.S.s0.i2 << 8) | ((unsigned)g.S.s1.i1 << 16) | ((unsigned)g.S.s1.i2 << 24);
            rec[8] = (unsigned)g.S.s2.i1 | ((unsigned)g.S.s2.i2 << 8) | ((unsigned)g.S.s3.i1 << 16) | ((unsigned)g.S.s3.i2 << 24);
            __threadfence_block();  // record (and this thread's slot reads) before the two flags
          }
          __syncwarp();
          if (finished) {
            st_vol(&ready[idx % kS16RingRecords], idx / kS16RingRecords + 1u);
            st_vol(&ctrl[cslot], kS16Free);
            state = kWait;
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      atomicAdd(&ring_ctl[2], 1u);
    }
  } else if (warp < CW + NC) {
    // =============================================== converter ===============================================
    constexpr int kPer = kS16Slots / NC;  // slots served by this warp
    constexpr int kWords = kPer / 32;
    const int first = (warp - CW) * kPer;
    const bool second = lane >= G1;  // this lane converts four vertices of body 2
    const bool act = lane < G1 + G2;
    const int lg = second ? lane - G1 : lane;  // 4-vertex group within the body
    const int src0 = second ? G1 : 0;          // lane holding the body's first vertex
    const uint32_t lane_off = kS16HeaderBytes + (uint32_t)(second ? NB1 * 48 : 0) + (uint32_t)(lg >> 1) * 48u + (uint32_t)(lg & 1) * 8u;
    uint4* pk = picks + (warp - CW) * P;  // this warp's batch: (slot, ticket, polytope index 1, polytope index 2)
    const unsigned lt = (1u << lane) - 1u;
    unsigned tk_next = 0, tk_end = 0;  // reserved ticket range (warp-uniform)
    unsigned raw = 0;                  // lane 0: result of the atomic that reserves the NEXT range, issued ahead of need
    bool pending = false;
    int exited = 0;
    for (;;) {
      // ---- collect up to P free slots of this warp's range: lane l watches slots first + l (+ 32, ...) ----
      unsigned fm[kWords];
      unsigned total = 0;
#pragma unroll
      for (int j = 0; j < kWords; ++j) {
        fm[j] = __ballot_sync(0xffffffffu, ld_vol(&ctrl[first + 32 * j + lane]) == kS16Free);
        total += __popc(fm[j]);
      }
      if (total == 0) {
        if (exited == kPer) break;
        __nanosleep(40);
        continue;
      }
      __threadfence_block();  // acquire: the owners' last slot reads precede the flags
      const unsigned cnt = total < (unsigned)P ? total : (unsigned)P;
      // tickets: ranks below `avail` come from the range in hand, the others from the next one
      const unsigned avail = tk_end - tk_next;
      unsigned nb = 0;
      if (cnt > avail) {
        if (!pending && lane == 0) raw = atomicAdd(ticket, kTicketChunk);
        nb = __shfl_sync(0xffffffffu, raw, 0);
        pending = false;
      }
      {
        unsigned before = 0;
#pragma unroll
        for (int j = 0; j < kWords; ++j) {
          const unsigned r = before + __popc(fm[j] & lt);
          if (((fm[j] >> lane) & 1u) && r < cnt) {
            const unsigned t = r < avail ? tk_next + r : nb + (r - avail);
            int a = (int)t, b = (int)t;
            if (IDX && t < n) {
              const CollisionPair pr = pairs[t];
              a = pr.idx1;
              b = pr.idx2;
            }
            pk[r] = make_uint4((unsigned)(32 * j + lane), t, (unsigned)a, (unsigned)b);
          }
          before += __popc(fm[j]);
        }
      }
      if (cnt > avail) {
        tk_next = nb + (cnt - avail);
        tk_end = nb + kTicketChunk;
      } else {
        tk_next += cnt;
      }
      __syncwarp();
      int sl[P];
      unsigned tk[P];
      int i1[P], i2[P];
      uint4 q[P][3];
#pragma unroll
      for (int k = 0; k < P; ++k) {
        const uint4 e = pk[k < (int)cnt ? k : 0];
        sl[k] = (int)e.x;
        tk[k] = e.y;
        i1[k] = (int)e.z;
        i2[k] = (int)e.w;
      }
      __syncwarp();  // the list is read before the next batch overwrites it
      // ---- all loads of the batch first: these registers are the bytes in flight ----
#pragma unroll
      for (int k = 0; k < P; ++k) {
        if (k < (int)cnt && tk[k] < n && act) {
          const float* src = (second ? coord2 + (size_t)i2[k] * (NV2 * 3) : coord1 + (size_t)i1[k] * (NV1 * 3)) + lg * 12;
          q[k][0] = ldg128(src);
          q[k][1] = ldg128(src + 4);
          q[k][2] = ldg128(src + 8);
        } else {
          q[k][0] = q[k][1] = q[k][2] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      if (!pending && tk_end - tk_next < (unsigned)P) {  // reserve the next ticket range behind the loads
        if (lane == 0) raw = atomicAdd(ticket, kTicketChunk);
        pending = true;
      }
      // ---- centre, scale, convert, store, publish ----
#pragma unroll
      for (int k = 0; k < P; ++k) {
        if (k >= (int)cnt) break;
        const int s = first + sl[k];
        if (tk[k] >= n) {  // out of work: the slot's owner may leave
          if (lane == 0) st_vol(&ctrl[s], kS16Exit);
          ++exited;
          continue;
        }
        float f[12];
        f[0] = __uint_as_float(q[k][0].x); f[1] = __uint_as_float(q[k][0].y); f[2] = __uint_as_float(q[k][0].z);
        f[3] = __uint_as_float(q[k][0].w); f[4] = __uint_as_float(q[k][1].x); f[5] = __uint_as_float(q[k][1].y);
        f[6] = __uint_as_float(q[k][1].z); f[7] = __uint_as_float(q[k][1].w); f[8] = __uint_as_float(q[k][2].x);
        f[9] = __uint_as_float(q[k][2].y); f[10] = __uint_as_float(q[k][2].z); f[11] = __uint_as_float(q[k][2].w);
        const float cx = __shfl_sync(0xffffffffu, f[0], src0), cy = __shfl_sync(0xffffffffu, f[1], src0),
                    cz = __shfl_sync(0xffffffffu, f[2], src0);
        float e[12];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          e[3 * i] = sub_rn(f[3 * i], cx);
          e[3 * i + 1] = sub_rn(f[3 * i + 1], cy);
          e[3 * i + 2] = sub_rn(f[3 * i + 2], cz);
        }
        float m = 0.0f;
#pragma unroll
        for (int i = 0; i < 12; ++i) m = fmaxf(m, fabsf(e[i]));
        const unsigned mb = act ? __float_as_uint(m) : 0u;  // non-negative floats order like their bit patterns
        const unsigned r1 = __reduce_max_sync(0xffffffffu, second ? 0u : mb);
        const unsigned r2 = __reduce_max_sync(0xffffffffu, second ? mb : 0u);
        const float mm = __uint_as_float(second ? r2 : r1);
        const bool ok = mm > 8.67361737988e-19f;  // 2^-60
        const float sc = ok ? __fdividef(kS16Scale, mm) : 0.0f;
        unsigned h[6];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const __half2 lo = __floats2half2_rn(mul_rn(e[c], sc), mul_rn(e[3 + c], sc));
          const __half2 hi = __floats2half2_rn(mul_rn(e[6 + c], sc), mul_rn(e[9 + c], sc));
          h[2 * c] = *reinterpret_cast<const unsigned*>(&lo);
          h[2 * c + 1] = *reinterpret_cast<const unsigned*>(&hi);
        }
        unsigned char* dst = slots + (size_t)s * sbytes;
        if (act) {
          *reinterpret_cast<uint2*>(dst + lane_off) = make_uint2(h[0], h[1]);
          *reinterpret_cast<uint2*>(dst + lane_off + 16) = make_uint2(h[2], h[3]);
          *reinterpret_cast<uint2*>(dst + lane_off + 32) = make_uint2(h[4], h[5]);
          if (lg == 0) {
            float* hd = reinterpret_cast<float*>(dst) + (second ? 6 : 0);
            hd[0] = f[0];
            hd[1] = f[1];
            hd[2] = f[2];
            hd[3] = ok ? fmaf(kS16WCentre * sc, fabsf(f[0]), kS16WConst) : 1e30f;
            hd[4] = ok ? fmaf(kS16WCentre * sc, fabsf(f[1]), kS16WConst) : 1e30f;
            hd[5] = ok ? fmaf(kS16WCentre * sc, fabsf(f[2]), kS16WConst) : 1e30f;
          }
        }
        if (lane == 0) {
          pair_of[s] = tk[k];
          if (IDX) {
            idx1_of[s] = i1[k];
            idx2_of[s] = i2[k];
          }
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) st_vol(&ctrl[s], kS16Ready);
      }
    }
  } else {
    // ================================================ finisher ===============================================
    float* mine = scratch + lane * kS16ScratchWords;
    unsigned cur = 0;
    for (;;) {
      const unsigned idx = cur + lane;
      const bool rdy = ld_vol(&ready[idx % kS16RingRecords]) == idx / kS16RingRecords + 1u;
      const unsigned m = __ballot_sync(0xffffffffu, rdy);
      const int c = (m == 0xffffffffu) ? 32 : (__ffs(~m) - 1);  // records ready in order from `cur`
      if (c == 0) {
        if (ld_vol(&ring_ctl[2]) == (unsigned)CW && ld_vol(&ring_ctl[0]) == cur) break;
        __nanosleep(100);
        continue;
      }
      __threadfence_block();
      if (lane < c) {
        const unsigned* rec = ring + (size_t)(idx % kS16RingRecords) * kS16RecWords;
        const unsigned pair = rec[0];
        GjkState<T> g;
        g.S.n = (int)rec[1];
        g.v = mk<T>(__uint_as_float(rec[2]), __uint_as_float(rec[3]), __uint_as_float(rec[4]));
        const float* b1 = coord1 + (size_t)rec[5] * (NV1 * 3);
        const float* b2 = coord2 + (size_t)rec[6] * (NV2 * 3);
        const unsigned w0 = rec[7], w1p = rec[8];
        const int vi[4][2] = {{(int)(w0 & 255u), (int)((w0 >> 8) & 255u)},
                              {(int)((w0 >> 16) & 255u), (int)(w0 >> 24)},
                              {(int)(w1p & 255u), (int)((w1p >> 8) & 255u)},
                              {(int)((w1p >> 16) & 255u), (int)(w1p >> 24)}};
        SV<T>* sv[4] = {&g.S.s0, &g.S.s1, &g.S.s2, &g.S.s3};
        float vt[4][6];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) {
            vt[k][cc] = __ldg(b1 + 3 * vi[k][0] + cc);
            vt[k][3 + cc] = __ldg(b2 + 3 * vi[k][1] + cc);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int cc = 0; cc < 6; ++cc) mine[6 * k + cc] = vt[k][cc];
          // the simplex point is the same fp32 subtraction the compute thread made when the vertex pair was added
          sv[k]->p = mk<T>(sub_rn(vt[k][0], vt[k][3]), sub_rn(vt[k][1], vt[k][4]), sub_rn(vt[k][2], vt[k][5]));
          // tag (bits 30..31): which scratch entry this slot came with -- survives the witness stage's slot shuffles
          sv[k]->i1 = (int)((unsigned)vi[k][0] | ((unsigned)k << 30));
          sv[k]->i2 = (int)((unsigned)vi[k][1] | ((unsigned)k << 30));
        }
        RecordFetch<T> fetch{mine};
        V3<T> w1, w2;
        gjk_witnesses(fetch, g.S, w1, w2);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sv[k]->i1 &= 0x3fffffff;
          sv[k]->i2 &= 0x3fffffff;
        }
        store_result(simplices + pair, distances + pair, g, w1, w2);
        if (normals) {  // fused EPA gate (EPA.c:369-373)
          const T dist = sqrt_rn(norm2(g.v));
          const bool collide = !(dist > Tol<T>::eps());
          if (!collide) {
            const V3<T> nr = normal_from_witnesses(w1, w2);
            T* o = normals + 3 * (size_t)pair;
            o[0] = nr.x;
            o[1] = nr.y;
            o[2] = nr.z;
          } else {
            epa_queue[atomicAdd(epa_count, 1)] = (int)pair;
          }
        }
      }
      __syncwarp();
      __threadfence_block();
      cur += (unsigned)c;
      if (lane == 0) st_vol(&ring_ctl[1], cur);
    }
  }
}

}  // namespace ogjk

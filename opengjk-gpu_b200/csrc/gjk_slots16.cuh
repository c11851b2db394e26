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
// Roles (one CTA per SM, 8 + 2 warps):
//   * compute warps: ONE THREAD PER PAIR as in gjk_slots.cuh (rotated loop, lane-uniform iteration gjk_substep_u).  There
//     is no loader: a warp refills its own slots.  At every half-trip it issues the 128-bit loads of up to K pairs for
//     its free slots (lane l: four consecutive vertices of one body, three loads) into a register ring, and at the next
//     half-trip all 32 lanes centre / scale / convert (packed FADD2 / FMUL2) those pairs and write their owners' fp16
//     slots in blocks of eight vertices  x0..x7 | y0..y7 | z0..z7  (three 128-bit shared loads per block in the scan).
//     The bytes in flight that the TMA version kept in idle slots are held in registers (8 warps x K pairs); producer
//     and consumer of a slot are the same warp, so __syncwarp() is the only synchronisation -- no flags, no mbarriers,
//     no polling, and ONE instruction stream for fetching, converting and iterating (a version with separate
//     converter warps was instruction-fetch bound: profiles/r2_experiments.txt);
//   * two finisher warps (four compute warps each) take 9-word records (pair, simplex size, v, vertex indices), re-read
//     the <= 8 source vertices from global memory/L2, rebuild the simplex points (the same fp32 subtraction), and run
//     the witness stage, the result stores and the fused EPA gate exactly as in gjk_slots.cuh.
#pragma once
#include <cuda_fp16.h>

#include "gjk_slots.cuh"

namespace ogjk {

constexpr int kS16Slots = 256;
constexpr int kS16ComputeWarps = 8;   // one slot per thread
constexpr int kS16RingRecords = 64;
constexpr int kS16RecWords = 9;       // pair | n | v.xyz | polytope index 1 | polytope index 2 | vertex indices (2 words)
constexpr int kS16Finishers = 2;      // finisher f serves compute warps 4 f .. 4 f + 3 through its own record ring
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
  // table | per finisher: ring control (16 B), ready flags, ring
  return kSlotTableBytes +
         (uint32_t)kS16Finishers * (16u + (uint32_t)kS16RingRecords * 4u + (((uint32_t)kS16RingRecords * kS16RecWords * 4u + 15u) & ~15u));
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

// the finisher's vertex fetch: the four vertex pairs of a record live in registers; the tag in bits 30..31 of the
// index picks one with selects
struct RegFetch16 {
  float v[4][6];
  OGJK_D V3<float> operator()(int body, int i) const {
    const unsigned k = (unsigned)i >> 30;
    const int o = 3 * body;
    const float x = k == 0u ? v[0][o] : k == 1u ? v[1][o] : k == 2u ? v[2][o] : v[3][o];
    const float y = k == 0u ? v[0][o + 1] : k == 1u ? v[1][o + 1] : k == 2u ? v[2][o + 1] : v[3][o + 1];
    const float z = k == 0u ? v[0][o + 2] : k == 1u ? v[1][o + 2] : k == 2u ? v[2][o + 2] : v[3][o + 2];
    return mk<float>(x, y, z);
  }
};
OGJK_D uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
OGJK_D u64 sub2(u64 a, u64 b) {  // two individually rounded fp32 subtractions in one issue slot (FADD2)
  u64 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
OGJK_D u64 mul2s(u64 a, float s) {  // both halves times s (FMUL2)
  u64 ss, r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(ss) : "f"(s));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(ss));
  return r;
}
OGJK_D unsigned atom_add_global(unsigned* p, unsigned v) {  // one ATOMG, without the compiler's warp-aggregation prologue
  unsigned r;
  asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory");
  return r;
}
OGJK_D float lo32(u64 a) { return __uint_as_float((unsigned)a); }
OGJK_D float hi32(u64 a) { return __uint_as_float((unsigned)(a >> 32)); }

// NB1 / NB2: blocks of eight vertices per body (vertex counts 8 NB1 and 8 NB2, 2 (NB1 + NB2) <= 32 converting lanes).
// IDX: pairs are gkCollisionPair records into one pool.  K: pairs a warp keeps in flight (register ring).
constexpr int kS16Threads = (kS16ComputeWarps + kS16Finishers) * 32;
template <int NB1, int NB2, bool IDX, int K>
__global__ void __launch_bounds__(kS16Threads)
gjk_slots16_kernel(const float* __restrict__ coord1, const float* __restrict__ coord2, SimplexT<float>* __restrict__ simplices,
                   float* __restrict__ distances, unsigned n, const uint16_t* __restrict__ utab_g,
                   unsigned* __restrict__ ticket, float* __restrict__ normals, int* __restrict__ epa_queue,
                   int* __restrict__ epa_count, const CollisionPair* __restrict__ pairs) {
  typedef float T;
  constexpr int CW = kS16ComputeWarps;
  constexpr int kThreads = kS16Threads;
  constexpr int NV1 = 8 * NB1, NV2 = 8 * NB2;
  constexpr int G1 = 2 * NB1, G2 = 2 * NB2;  // converting lanes per body (four vertices each)
  static_assert(G1 + G2 <= 32, "a pair must fit the 32 lanes of a warp");
  constexpr uint32_t sbytes = s16_slot_bytes(NB1, NB2);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* sp = smem_raw;
  uint16_t* utab = reinterpret_cast<uint16_t*>(sp);
  sp += kSlotTableBytes;
  // record rings: compute warps 0..3 feed finisher 0, warps 4..7 finisher 1
  constexpr uint32_t kRingBytes = 16u + kS16RingRecords * 4u + ((kS16RingRecords * kS16RecWords * 4u + 15u) & ~15u);
  const int fin_id = warp < CW ? warp / (CW / kS16Finishers) : warp - CW;
  unsigned* ring_ctl = reinterpret_cast<unsigned*>(sp + fin_id * kRingBytes);  // [0] tail (reserved), [1] head (consumed), [2] exited warps
  unsigned* ready = ring_ctl + 4;
  unsigned* ring = ready + kS16RingRecords;
  unsigned char* slots = smem_raw + s16_fixed_bytes();

  for (int i = tid; i < kUnifiedSize / 2; i += kThreads)
    reinterpret_cast<uint32_t*>(utab)[i] = __ldg(reinterpret_cast<const uint32_t*>(utab_g) + i);
  for (int i = tid; i < (int)(kS16Finishers * kRingBytes / 4u); i += kThreads) reinterpret_cast<unsigned*>(sp)[i] = 0;
  __syncthreads();

  if (warp < CW) {
    // =================================== compute + conversion (one code stream) ===================================
    // Lane l owns slot 32 warp + l for every phase of its pair; the warp as a whole fetches and converts the pairs of
    // its own free slots: at the top of a half-trip a pair's two fp32 vertex sets are requested with three 128-bit loads
    // per lane (lane l: four consecutive vertices) into K register sets, and at the bottom -- behind the half-trip's
    // arithmetic, which is what hides the latency -- they are centred / scaled / converted / stored into the owner's
    // slot.  K pairs per warp are in flight (the bytes in flight that the TMA version kept in idle slots are held in
    // registers), and no flag, barrier or polling warp is involved: producer and consumer are the same warp,
    // __syncwarp() orders the slot stores.  (Loads and conversion sit in ONE loop iteration on purpose: across the
    // back-edge ptxas waits for every outstanding load at the loop top, profiles/r2_experiments.txt.)
    const unsigned char* sbase = slots + (size_t)tid * sbytes;
    const float* hdr = reinterpret_cast<const float*>(sbase);
    const uint4* blk1 = reinterpret_cast<const uint4*>(sbase + kS16HeaderBytes);
    const uint4* blk2 = blk1 + 3 * NB1;
    unsigned char* const wslots = slots + (size_t)(warp * 32) * sbytes;  // this warp's 32 slots
    // conversion role of this lane
    const bool second = lane >= G1;  // four vertices of body 2
    const bool act = lane < G1 + G2;
    const int lg = second ? lane - G1 : lane;  // 4-vertex group within the body
    const int src0 = second ? G1 : 0;          // lane holding the body's first vertex
    const uint32_t lane_off = kS16HeaderBytes + (uint32_t)(second ? NB1 * 48 : 0) + (uint32_t)(lg >> 1) * 48u + (uint32_t)(lg & 1) * 8u;
    const float* const lane_src = (second ? coord2 : coord1) + (act ? lg : 0) * 12;  // idle lanes re-read group 0
    constexpr int kBodyFloats1 = NV1 * 3, kBodyFloats2 = NV2 * 3;
    // tickets (warp-uniform): the range in use [tk_next, tk_end) and a spare [sp_next, sp_end) requested with one atomic
    // and collected a half-trip later, together with -- for indexed batches -- the range's pair records (one per lane)
    unsigned tk_next = 0, tk_end = 0, tk_base = 0, sp_next = 0, sp_end = 0;
    unsigned raw = 0;  // lane 0: result of the atomic in flight
    bool pending = false;
    int rec1 = 0, rec2 = 0, srec1 = 0, srec2 = 0;
    // the register ring
    ulonglong2 q[K][3];  // x0 y0 | z0 x1 || y1 z1 | x2 y2 || z2 x3 | y3 z3
    int sl[K];
    unsigned tk[K];
    int ri1[K], ri2[K];
#pragma unroll
    for (int k = 0; k < K; ++k) sl[k] = -1;
    unsigned freemask = 0xffffffffu;  // slots waiting for a pair (warp-uniform)
    unsigned exitmask = 0;            // slots that were told there is no more work
    enum { kIdle = 0, kFresh = 1, kRun = 2 };
    int state = kIdle;
    unsigned pair = 0;
    int pi1 = 0, pi2 = 0;
    const float* g1 = coord1;  // this pair's fp32 vertices in global memory
    const float* g2 = coord2;
    GjkState<T> g;
    bool need_sub = false;
    // A trip is two half-trips: [sub-algorithm step of the iteration begun last trip | pre-scans + verification + the two
    // exit pre-tests of the next iteration], each preceded by [request the vertices for the free slots] and followed by
    // [retire the lanes that finished | convert what was requested].  The arithmetic per pair is the sequence of
    // gjk_advance_u.
    for (int phase = 0;; phase ^= 1) {
      // ---- tickets: collect the spare range requested a half-trip ago, or request one ----
      if (sp_next == sp_end) {
        if (!pending) {
          if (lane == 0) raw = atom_add_global(ticket, kTicketChunk);
          pending = true;
        } else {
          sp_next = __shfl_sync(0xffffffffu, raw, 0);
          sp_end = sp_next + kTicketChunk;
          pending = false;
          if (IDX) {
            srec1 = srec2 = 0;
            if (sp_next + lane < n) {
              const CollisionPair pr = pairs[sp_next + lane];
              srec1 = pr.idx1;
              srec2 = pr.idx2;
            }
          }
        }
      }
      if (tk_next == tk_end && sp_next != sp_end) {  // the range in use is exhausted: the spare takes over
        tk_base = tk_next = sp_next;
        tk_end = sp_end;
        sp_next = sp_end;
        rec1 = srec1;
        rec2 = srec2;
      }
      // ---- fetch for the free slots: these registers are the bytes in flight ----
      // (the loads are issued unconditionally -- from pair 0 when an entry stays empty -- so that the ring registers are
      //  written by the loads alone: a conditional assignment makes the compiler wait for the load right here)
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const bool have = freemask != 0u && tk_next != tk_end;  // warp-uniform
        const int s = have ? __ffs((int)freemask) - 1 : 0;
        const unsigned t = have ? tk_next : 0u;
        const bool take = have && t < n;
        if (have) {
          freemask &= freemask - 1u;
          ++tk_next;
          if (!take) exitmask |= 1u << s;  // out of work
        }
        int a = (int)t, b = (int)t;
        if (IDX) {
          a = __shfl_sync(0xffffffffu, rec1, (int)((t - tk_base) & 31u));
          b = __shfl_sync(0xffffffffu, rec2, (int)((t - tk_base) & 31u));
        }
        if (!take) a = b = 0;
        sl[k] = take ? s : -1;
        tk[k] = t;
        ri1[k] = a;
        ri2[k] = b;
        const float* src = lane_src + (second ? (size_t)b * kBodyFloats2 : (size_t)a * kBodyFloats1);
        q[k][0] = __ldg(reinterpret_cast<const ulonglong2*>(src));
        q[k][1] = __ldg(reinterpret_cast<const ulonglong2*>(src + 4));
        q[k][2] = __ldg(reinterpret_cast<const ulonglong2*>(src + 8));
      }
      bool finished = false;
      if (phase == 0) {
        if (state == kRun && need_sub) finished = gjk_substep_u(g, utab);
        need_sub = false;
      } else {
        if (state == kFresh) {  // converted at the last half-trip
          g1 = coord1 + (size_t)pi1 * (NV1 * 3);
          g2 = coord2 + (size_t)pi2 * (NV2 * 3);
          gjk_init(g, mk<T>(hdr[0], hdr[1], hdr[2]), mk<T>(hdr[6], hdr[7], hdr[8]));
          state = kRun;
        }
        if (state == kRun) {
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
            V3<T> p = g.sup1, qq = g.sup2;
            if (i1 >= 0) {
              const float* a = g1 + 3 * i1;
              p = mk<T>(__ldg(a), __ldg(a + 1), __ldg(a + 2));
            }
            if (i2 >= 0) {
              const float* a = g2 + 3 * i2;
              qq = mk<T>(__ldg(a), __ldg(a + 1), __ldg(a + 2));
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
              const T dd = dot(qq, g.v);
              if (dd > best2) {
                best2 = dd;
                g.sup2 = qq;
                g.idx2 = i2;
              }
            }
          }
          finished = gjk_converged_u(g);
          need_sub = !finished;
        }
      }
      // ---- retire: 9-word record for the finisher; the slot joins the free list ----
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
            __threadfence_block();  // record before its flag
            state = kIdle;
          }
          __syncwarp();
          if (finished) st_vol(&ready[idx % kS16RingRecords], idx / kS16RingRecords + 1u);
          freemask |= fin;
        }
      }
      // ---- convert the pairs fetched at the top of this half-trip into their owners' slots ----
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (sl[k] < 0) continue;  // warp-uniform
        const int s = sl[k];
        const u64 p0 = q[k][0].x, p1 = q[k][0].y, p2 = q[k][1].x, p3 = q[k][1].y, p4 = q[k][2].x, p5 = q[k][2].y;
        const float f0 = lo32(p0), f1 = hi32(p0), f2 = lo32(p1);
        const float cx = __shfl_sync(0xffffffffu, f0, src0), cy = __shfl_sync(0xffffffffu, f1, src0),
                    cz = __shfl_sync(0xffffffffu, f2, src0);
        const u64 cxy = pack2(cx, cy), czx = pack2(cz, cx), cyz = pack2(cy, cz);
        const u64 e0 = sub2(p0, cxy), e1 = sub2(p1, czx), e2 = sub2(p2, cyz), e3 = sub2(p3, cxy), e4 = sub2(p4, czx),
                  e5 = sub2(p5, cyz);
        float m = fmaxf(fabsf(lo32(e0)), fabsf(hi32(e0)));
        m = fmaxf(m, fmaxf(fabsf(lo32(e1)), fabsf(hi32(e1))));
        m = fmaxf(m, fmaxf(fabsf(lo32(e2)), fabsf(hi32(e2))));
        m = fmaxf(m, fmaxf(fabsf(lo32(e3)), fabsf(hi32(e3))));
        m = fmaxf(m, fmaxf(fabsf(lo32(e4)), fabsf(hi32(e4))));
        m = fmaxf(m, fmaxf(fabsf(lo32(e5)), fabsf(hi32(e5))));
        const unsigned mb = act ? __float_as_uint(m) : 0u;  // non-negative floats order like their bit patterns
        const unsigned r1 = __reduce_max_sync(0xffffffffu, second ? 0u : mb);
        const unsigned r2 = __reduce_max_sync(0xffffffffu, second ? mb : 0u);
        const float mm = __uint_as_float(second ? r2 : r1);
        const bool ok = mm > 8.67361737988e-19f;  // 2^-60
        const float sc = ok ? __fdividef(kS16Scale, mm) : 0.0f;
        const u64 g0 = mul2s(e0, sc), g1v = mul2s(e1, sc), g2v = mul2s(e2, sc), g3 = mul2s(e3, sc), g4 = mul2s(e4, sc),
                  g5 = mul2s(e5, sc);
        // vertices: 0 = (g0.lo, g0.hi, g1.lo)  1 = (g1.hi, g2.lo, g2.hi)  2 = (g3.lo, g3.hi, g4.lo)  3 = (g4.hi, g5.lo, g5.hi)
        const __half2 x01 = __floats2half2_rn(lo32(g0), hi32(g1v)), x23 = __floats2half2_rn(lo32(g3), hi32(g4));
        const __half2 y01 = __floats2half2_rn(hi32(g0), lo32(g2v)), y23 = __floats2half2_rn(hi32(g3), lo32(g5));
        const __half2 z01 = __floats2half2_rn(lo32(g1v), hi32(g2v)), z23 = __floats2half2_rn(lo32(g4), hi32(g5));
        unsigned char* dst = wslots + (size_t)s * sbytes;
        if (act) {
          *reinterpret_cast<uint2*>(dst + lane_off) =
              make_uint2(*reinterpret_cast<const unsigned*>(&x01), *reinterpret_cast<const unsigned*>(&x23));
          *reinterpret_cast<uint2*>(dst + lane_off + 16) =
              make_uint2(*reinterpret_cast<const unsigned*>(&y01), *reinterpret_cast<const unsigned*>(&y23));
          *reinterpret_cast<uint2*>(dst + lane_off + 32) =
              make_uint2(*reinterpret_cast<const unsigned*>(&z01), *reinterpret_cast<const unsigned*>(&z23));
        }
        {  // header: c0 and W of this lane's body, stored by the lane that holds the first vertex
          const float ws = kS16WCentre * sc;
          const float w0 = ok ? fmaf(ws, fabsf(f0), kS16WConst) : 1e30f, w1 = ok ? fmaf(ws, fabsf(f1), kS16WConst) : 1e30f,
                      w2 = ok ? fmaf(ws, fabsf(f2), kS16WConst) : 1e30f;
          if (act && lg == 0) {
            float2* hd = reinterpret_cast<float2*>(dst + (second ? 24 : 0));
            hd[0] = make_float2(f0, f1);
            hd[1] = make_float2(f2, w0);
            hd[2] = make_float2(w1, w2);
          }
        }
        if (lane == s) {  // the owner: its pair starts at the next scan half-trip
          pair = tk[k];
          pi1 = ri1[k];
          pi2 = ri2[k];
          state = kFresh;
        }
        sl[k] = -1;
      }
      __syncwarp();  // the slots' stores are ordered before their owners' loads
      if (exitmask == 0xffffffffu) break;  // every slot has been told: nothing in flight, nothing running
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      atomicAdd(&ring_ctl[2], 1u);
    }
  } else {
    // ================================================ finisher ===============================================
    unsigned cur = 0;
    for (;;) {
      const unsigned idx = cur + lane;
      const bool rdy = ld_vol(&ready[idx % kS16RingRecords]) == idx / kS16RingRecords + 1u;
      const unsigned m = __ballot_sync(0xffffffffu, rdy);
      const int c = (m == 0xffffffffu) ? 32 : (__ffs(~m) - 1);  // records ready in order from `cur`
      if (c == 0) {
        if (ld_vol(&ring_ctl[2]) == (unsigned)(CW / kS16Finishers) && ld_vol(&ring_ctl[0]) == cur) break;
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
          // the simplex point is the same fp32 subtraction the compute thread made when the vertex pair was added
          sv[k]->p = mk<T>(sub_rn(vt[k][0], vt[k][3]), sub_rn(vt[k][1], vt[k][4]), sub_rn(vt[k][2], vt[k][5]));
          // tag (bits 30..31): which fetched vertex pair this slot came with -- survives the witness stage's slot shuffles
          sv[k]->i1 = (int)((unsigned)vi[k][0] | ((unsigned)k << 30));
          sv[k]->i2 = (int)((unsigned)vi[k][1] | ((unsigned)k << 30));
        }
        const RegFetch16 fetch{{{vt[0][0], vt[0][1], vt[0][2], vt[0][3], vt[0][4], vt[0][5]},
                                {vt[1][0], vt[1][1], vt[1][2], vt[1][3], vt[1][4], vt[1][5]},
                                {vt[2][0], vt[2][1], vt[2][2], vt[2][3], vt[2][4], vt[2][5]},
                                {vt[3][0], vt[3][1], vt[3][2], vt[3][3], vt[3][4], vt[3][5]}}};
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

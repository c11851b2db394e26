// gjk_slots.cuh -- persistent "slot" GJK kernel for uniform fp32 batches: the Blackwell-native hot path.
//
// ONE THREAD OWNS ONE PAIR for every phase (support scans, exit tests, sub-algorithm, witnesses), so no instruction
// is executed redundantly by cooperating lanes, and the kernel is persistent so no lane idles while others iterate:
//   * each thread has a private shared-memory slot holding its pair's two vertex sets exactly as they lie in HBM
//     (xyz interleaved).  A finished thread takes the next pair index from a global ticket and refills its slot with
//     two TMA bulk copies (cp.async.bulk global->shared, completion on the slot's own mbarrier); while the copy is
//     in flight the other 31 lanes keep iterating.  This is the work queue that rebalances pairs whose iteration
//     counts diverge (1..25 iterations, mean 3.8 at 64 vertices).  Tickets are drawn per warp in chunks of 32 (one
//     atomic per ~8 warp iterations instead of one per iteration);
//   * slot stride is an odd multiple of 16 bytes, so the 128-bit shared loads of the 32 lanes of a warp (each in
//     its own slot) are bank-conflict free;
//   * the support scan walks the slot four vertices (three 128-bit loads) at a time with packed FMUL2 products and
//     scalar adds (ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false, which would break
//     bit-parity with the reference's unfused arithmetic, so only the multiplies are packed), keeps the running
//     maximum and the index of the winning 4-vertex block, and recovers the exact lowest winning index from that
//     one block afterwards (SURVEY.md Appendix A.2 tie-break).  The loads are software-pipelined one block ahead:
//     with one warp per scheduler (a 64+64-vertex slot is 1.5 KB, so 128 threads fill an SM's shared memory) there
//     is no other warp to hide the 29-cycle LDS latency;
//   * everything after the scans is the LANE-UNIFORM iteration gjk_advance_u (gjk_core.cuh): threads holding 2-, 3-
//     and 4-point simplices execute one instruction stream driven by one 16-bit table (copied to shared memory),
//     instead of three divergent sub-algorithm paths serialised by the SIMT hardware (profiles/r1c_gjk_slots_v1.txt:
//     the v1 kernel spent 72 % of its issue slots in those paths at ~5 active lanes);
//   * witness vertices are fetched from the slot, so global memory is touched once per vertex.
// HBM traffic is the algorithmic minimum: every vertex byte is read once (by TMA), every result byte written once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "epa_kernel.cuh"
#include "gjk_core.cuh"
#include "gjk_generic.cuh"
#include "gjk_tables.h"
#include "ogjk_types.h"

namespace ogjk {

// ---- mbarrier / TMA bulk-copy primitives (PTX ISA: mbarrier, cp.async.bulk) --------------------------------------
OGJK_D uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
OGJK_D void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
OGJK_D void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
OGJK_D bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
OGJK_D void tma_bulk_load(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src_gmem), "r"(bytes), "r"(bar)
               : "memory");
}
OGJK_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- packed products ----------------------------------------------------------------------------------------------
typedef unsigned long long u64;
OGJK_D u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
OGJK_D void mul2(u64 a, u64 b, float& lo, float& hi) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
}
// the search direction paired up the way three consecutive 16-byte chunks of xyz-interleaved vertices need it.
// `zero` is a kernel argument that is always 0: xor-ing it into the second copy of each component keeps ptxas from
// proving that the six halves are three values, so it keeps three 64-bit pairs live instead of re-assembling a
// pair with a MOV in front of every second FMUL2.
struct DirPack {
  u64 xy, zx, yz;
};
OGJK_D DirPack make_dir(const V3<float>& d, unsigned zero) {
  const float x2 = __uint_as_float(__float_as_uint(d.x) ^ zero);
  const float y2 = __uint_as_float(__float_as_uint(d.y) ^ zero);
  const float z2 = __uint_as_float(__float_as_uint(d.z) ^ zero);
  DirPack p;
  p.xy = pack2(d.x, d.y);
  p.zx = pack2(d.z, x2);
  p.yz = pack2(y2, z2);
  return p;
}
// one block = four vertices = 48 bytes:  x0 y0 | z0 x1 || y1 z1 | x2 y2 || z2 x3 | y3 z3
struct Block4 {
  ulonglong2 a, b, c;
};
OGJK_D Block4 load_block(const ulonglong2* chunk, int g) {
  Block4 k;
  k.a = chunk[3 * g];
  k.b = chunk[3 * g + 1];
  k.c = chunk[3 * g + 2];
  return k;
}
OGJK_D void dots4(const Block4& k, const DirPack& D, float (&dd)[4]) {
  float p0l, p0h, p1l, p1h, p2l, p2h, p3l, p3h, p4l, p4h, p5l, p5h;
  mul2(k.a.x, D.xy, p0l, p0h);
  mul2(k.a.y, D.zx, p1l, p1h);
  mul2(k.b.x, D.yz, p2l, p2h);
  mul2(k.b.y, D.xy, p3l, p3h);
  mul2(k.c.x, D.zx, p4l, p4h);
  mul2(k.c.y, D.yz, p5l, p5h);
  dd[0] = add_rn(add_rn(p0l, p0h), p1l);
  dd[1] = add_rn(add_rn(p1h, p2l), p2h);
  dd[2] = add_rn(add_rn(p3l, p3h), p4l);
  dd[3] = add_rn(add_rn(p4h, p5l), p5h);
}
OGJK_D float max4(const float (&dd)[4]) { return fmaxf(fmaxf(dd[0], dd[1]), fmaxf(dd[2], dd[3])); }

// Support search of one body by ONE thread over its shared-memory slot.  `body` points at nv*3 floats, nv % 4 == 0.
// Two blocks per trip with the next trip's loads issued first.  The look-ahead of the last trip reads up to 96
// bytes past the body -- the other body, the next slot or the pad at the end of the allocation -- and is discarded.
OGJK_D void support_slot(const float* body, int nv, const V3<float>& d, unsigned zero, V3<float>& sup, int& sup_idx) {
  const ulonglong2* chunk = reinterpret_cast<const ulonglong2*>(body);
  const DirPack D = make_dir(d, zero);
  float best = -INFINITY;
  int bg = 0;
  const int groups = nv >> 2;
  const int pairs = groups >> 1;
  Block4 k0 = load_block(chunk, 0);
  Block4 k1 = load_block(chunk, 1);
#pragma unroll 2
  for (int t = 0; t < pairs; ++t) {
    const Block4 n0 = load_block(chunk, 2 * t + 2);
    const Block4 n1 = load_block(chunk, 2 * t + 3);
    float da[4], db[4];
    dots4(k0, D, da);
    dots4(k1, D, db);
    const float ma = max4(da), mb = max4(db);
    if (ma > best) {  // strict: the earliest block holding the maximum wins
      best = ma;
      bg = 2 * t;
    }
    if (mb > best) {
      best = mb;
      bg = 2 * t + 1;
    }
    k0 = n0;
    k1 = n1;
  }
  if (groups & 1) {  // k0 holds the last block
    float da[4];
    dots4(k0, D, da);
    const float ma = max4(da);
    if (ma > best) {
      best = ma;
      bg = groups - 1;
    }
  }
  if (best > dot(sup, d)) {
    float dd[4];
    dots4(load_block(chunk, bg), D, dd);
    int k = 3;
    if (dd[2] == best) k = 2;
    if (dd[1] == best) k = 1;
    if (dd[0] == best) k = 0;
    const int idx = 4 * bg + k;
    sup = mk<float>(body[3 * idx], body[3 * idx + 1], body[3 * idx + 2]);
    sup_idx = idx;
  }
}

// Both bodies of a pair in one loop (equal vertex counts): two independent scan streams interleaved, so that the
// single warp a scheduler has (64+64-vertex slots) finds an independent instruction more often.
OGJK_D void recover_support(const float* body, const ulonglong2* chunk, int bg, float best, const DirPack& D,
                            const V3<float>& d, V3<float>& sup, int& sup_idx) {
  if (best > dot(sup, d)) {
    float dd[4];
    dots4(load_block(chunk, bg), D, dd);
    int k = 3;
    if (dd[2] == best) k = 2;
    if (dd[1] == best) k = 1;
    if (dd[0] == best) k = 0;
    const int idx = 4 * bg + k;
    sup = mk<float>(body[3 * idx], body[3 * idx + 1], body[3 * idx + 2]);
    sup_idx = idx;
  }
}
// (max, block) update of both bodies for blocks g and g + 1
OGJK_D void scan_two_blocks(const Block4& a0, const Block4& a1, const Block4& e0, const Block4& e1, const DirPack& D1,
                            const DirPack& D2, int g, float& best1, int& bg1, float& best2, int& bg2) {
  float da[4], db[4], dc[4], de[4];
  dots4(a0, D1, da);
  dots4(e0, D2, dc);
  dots4(a1, D1, db);
  dots4(e1, D2, de);
  const float ma = max4(da), mc = max4(dc), mb = max4(db), me = max4(de);
  if (ma > best1) {
    best1 = ma;
    bg1 = g;
  }
  if (mc > best2) {
    best2 = mc;
    bg2 = g;
  }
  if (mb > best1) {
    best1 = mb;
    bg1 = g + 1;
  }
  if (me > best2) {
    best2 = me;
    bg2 = g + 1;
  }
}
// Two register sets: while set A (blocks 2t, 2t+1 of both bodies) is evaluated the loads of set B (blocks 2t+2, 2t+3)
// are in flight and vice versa.  (With ONE set and `current = next` at the end of the trip the compiler lets `next` share
// the registers of `current`, so a load can only issue after the last use of the block it replaces: nine of the twelve
// loads of a trip ended up in its last 22 instructions and the first multiply of the next trip waited for them.)
OGJK_D void support_slots_both(const float* b1, const float* b2, int nv, const V3<float>& v, unsigned zero,
                               V3<float>& sup1, int& idx1, V3<float>& sup2, int& idx2) {
  const ulonglong2* c1 = reinterpret_cast<const ulonglong2*>(b1);
  const ulonglong2* c2 = reinterpret_cast<const ulonglong2*>(b2);
  const V3<float> nvv = vneg(v);
  const DirPack D1 = make_dir(nvv, zero), D2 = make_dir(v, zero);
  float best1 = -INFINITY, best2 = -INFINITY;
  int bg1 = 0, bg2 = 0;
  const int groups = nv >> 2;
  const int pairs = groups >> 1;
  Block4 a0 = load_block(c1, 0), a1 = load_block(c1, 1);
  Block4 e0 = load_block(c2, 0), e1 = load_block(c2, 1);
  int t = 0;
#pragma unroll 1
  for (; t + 2 <= pairs; t += 2) {
    const Block4 p0 = load_block(c1, 2 * t + 2), p1 = load_block(c1, 2 * t + 3);
    const Block4 r0 = load_block(c2, 2 * t + 2), r1 = load_block(c2, 2 * t + 3);
    scan_two_blocks(a0, a1, e0, e1, D1, D2, 2 * t, best1, bg1, best2, bg2);
    a0 = load_block(c1, 2 * t + 4);
    a1 = load_block(c1, 2 * t + 5);
    e0 = load_block(c2, 2 * t + 4);
    e1 = load_block(c2, 2 * t + 5);
    scan_two_blocks(p0, p1, r0, r1, D1, D2, 2 * t + 2, best1, bg1, best2, bg2);
  }
  if (t < pairs) {  // an odd number of block pairs: set A holds the last one
    const Block4 n0 = load_block(c1, 2 * t + 2), m0 = load_block(c2, 2 * t + 2);
    scan_two_blocks(a0, a1, e0, e1, D1, D2, 2 * t, best1, bg1, best2, bg2);
    a0 = n0;
    e0 = m0;
  }
  if (groups & 1) {  // a0 / e0 hold the last block
    float da[4], dc[4];
    dots4(a0, D1, da);
    dots4(e0, D2, dc);
    const float ma = max4(da), mc = max4(dc);
    if (ma > best1) {
      best1 = ma;
      bg1 = groups - 1;
    }
    if (mc > best2) {
      best2 = mc;
      bg2 = groups - 1;
    }
  }
  recover_support(b1, c1, bg1, best1, D1, nvv, sup1, idx1);
  recover_support(b2, c2, bg2, best2, D2, v, sup2, idx2);
}

// ---- SoA-4 packed bodies (indexed pools re-packed on the device: x0..x3 | y0..y3 | z0..z3 per block of four vertices) --
// With the coordinates transposed the two additions of a dot product can be packed as well: FFMA2(p, 1, q) rounds p + q
// once, exactly like the reference's unfused add (`one` is an opaque 1.0f so that ptxas can neither fold the multiply by
// one nor contract it with the preceding FMUL2).  Ten issue slots per four vertices instead of fourteen.
OGJK_D u64 mul2b(u64 a, float s) {  // both halves times s
  u64 ss, r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(ss) : "f"(s));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(ss));
  return r;
}
OGJK_D u64 add2_via_fma(u64 p, u64 one2, u64 q) {  // (p.lo + q.lo, p.hi + q.hi), each rounded once
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(p), "l"(one2), "l"(q));
  return r;
}
OGJK_D void dots4_pk(const Block4& k, const V3<float>& d, u64 one2, float (&dd)[4]) {
  const u64 s01 = add2_via_fma(mul2b(k.c.x, d.z), one2, add2_via_fma(mul2b(k.b.x, d.y), one2, mul2b(k.a.x, d.x)));
  const u64 s23 = add2_via_fma(mul2b(k.c.y, d.z), one2, add2_via_fma(mul2b(k.b.y, d.y), one2, mul2b(k.a.y, d.x)));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(dd[0]), "=f"(dd[1]) : "l"(s01));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(dd[2]), "=f"(dd[3]) : "l"(s23));
}
OGJK_D V3<float> packed_vertex(const float* body, int i) {
  const float* b = body + 12 * (i >> 2) + (i & 3);
  return mk<float>(b[0], b[4], b[8]);
}
OGJK_D void recover_support_pk(const float* body, const ulonglong2* chunk, int bg, float best, const V3<float>& d, u64 one2,
                               V3<float>& sup, int& sup_idx) {
  if (best > dot(sup, d)) {
    float dd[4];
    dots4_pk(load_block(chunk, bg), d, one2, dd);
    int k = 3;
    if (dd[2] == best) k = 2;
    if (dd[1] == best) k = 1;
    if (dd[0] == best) k = 0;
    const int idx = 4 * bg + k;
    sup = packed_vertex(body, idx);
    sup_idx = idx;
  }
}
// both bodies of a pair in one loop, as support_slots_both
OGJK_D void support_slots_both_pk(const float* b1, const float* b2, int nv, const V3<float>& v, unsigned zero,
                                  V3<float>& sup1, int& idx1, V3<float>& sup2, int& idx2) {
  const ulonglong2* c1 = reinterpret_cast<const ulonglong2*>(b1);
  const ulonglong2* c2 = reinterpret_cast<const ulonglong2*>(b2);
  const V3<float> nvv = vneg(v);
  const float one = __uint_as_float(0x3f800000u ^ zero);
  const u64 one2 = pack2(one, one);
  float best1 = -INFINITY, best2 = -INFINITY;
  int bg1 = 0, bg2 = 0;
  const int groups = nv >> 2;
  Block4 a0 = load_block(c1, 0), e0 = load_block(c2, 0);
#pragma unroll 2
  for (int t = 0; t < groups; ++t) {
    const Block4 na0 = load_block(c1, t + 1), ne0 = load_block(c2, t + 1);  // look-ahead; past the end it is discarded
    float da[4], dc[4];
    dots4_pk(a0, nvv, one2, da);
    dots4_pk(e0, v, one2, dc);
    const float ma = max4(da), mc = max4(dc);
    if (ma > best1) {  // strict: the earliest block holding the maximum wins
      best1 = ma;
      bg1 = t;
    }
    if (mc > best2) {
      best2 = mc;
      bg2 = t;
    }
    a0 = na0;
    e0 = ne0;
  }
  recover_support_pk(b1, c1, bg1, best1, nvv, one2, sup1, idx1);
  recover_support_pk(b2, c2, bg2, best2, v, one2, sup2, idx2);
}

// ---- per-warp ticket feed ------------------------------------------------------------------------------------------
// Pair indices are handed out through one global counter, 32 at a time per warp.  A warp holds the chunk it is
// consuming and prepares the NEXT one in the background, in three steps spread over successive calls so that neither
// the atomic's round trip nor -- for indexed batches (pairs = gkCollisionPair records into one polytope pool,
// reference openGJK.cu:1451-1476) -- the load of the chunk's records ever sits between a finished pair and its
// refill: (1) a quarter into the current chunk lane 0 issues the atomic, (2) once 20 tickets are gone its result is
// broadcast and lane l loads the record of ticket nbase + l into registers, (3) when the chunk runs out the
// registers become the current chunk.  `take` serves the lanes that raise `want` in ticket order; a request that
// straddles two chunks is served in two passes.
constexpr unsigned kTicketChunk = 32;
struct TicketFeed {
  unsigned base, next;  // current chunk [base, base + 32), next ticket to hand out
  unsigned nbase, raw;  // the chunk being prepared: its first ticket (valid at stage 2), lane 0's atomic result
  int stage;            // 0 nothing requested, 1 atomic issued, 2 nbase + records requested
  int c1, c2, n1, n2;   // this lane's record of the current / next chunk (indexed batches)

  OGJK_D void request(unsigned* ticket, int lane) {
    if (lane == 0) raw = atomicAdd(ticket, kTicketChunk);
    stage = 1;
  }
  OGJK_D void fetch(const CollisionPair* __restrict__ pairs, unsigned n, int lane) {
    nbase = __shfl_sync(0xffffffffu, raw, 0);
    n1 = n2 = 0;
    if (pairs && nbase + lane < n) {
      const CollisionPair pr = pairs[nbase + lane];
      n1 = pr.idx1;
      n2 = pr.idx2;
    }
    stage = 2;
  }
  OGJK_D void swap_in(unsigned* ticket, const CollisionPair* __restrict__ pairs, unsigned n, int lane) {
    if (stage == 0) request(ticket, lane);
    if (stage == 1) fetch(pairs, n, lane);
    base = next = nbase;
    c1 = n1;
    c2 = n2;
    stage = 0;
  }
  OGJK_D void init(unsigned* ticket, const CollisionPair* __restrict__ pairs, unsigned n, int lane) {
    base = next = nbase = raw = 0;
    c1 = c2 = n1 = n2 = 0;
    stage = 0;
    swap_in(ticket, pairs, n, lane);
  }
  // One pass: serves up to `avail` requesting lanes.  Returns true for a served lane and sets its ticket `t` and
  // polytope indices (i1, i2) -- (t, t) for dense batches.  Call from all 32 lanes; repeat while lanes still want.
  OGJK_D bool take(bool want, unsigned* ticket, const CollisionPair* __restrict__ pairs, unsigned n, int lane,
                   unsigned& t, int& i1, int& i2) {
    const unsigned wm = __ballot_sync(0xffffffffu, want);
    if (!wm) return false;
    const unsigned used = next - base;
    if (stage == 0 && used >= 8) request(ticket, lane);
    else if (stage == 1 && used >= 20) fetch(pairs, n, lane);
    if (used == kTicketChunk) swap_in(ticket, pairs, n, lane);
    const unsigned avail = base + kTicketChunk - next, cnt = __popc(wm);
    const unsigned r = __popc(wm & ((1u << lane) - 1u));
    const bool served = want && r < avail;
    const unsigned mine = next + r;
    const int src = (int)((mine - base) & 31u);
    const int r1 = __shfl_sync(0xffffffffu, c1, src), r2 = __shfl_sync(0xffffffffu, c2, src);
    if (served) {
      t = mine;
      i1 = pairs ? r1 : (int)mine;
      i2 = pairs ? r2 : (int)mine;
    }
    next += cnt < avail ? cnt : avail;
    return served;
  }
};

// ---- fp64 slots ----------------------------------------------------------------------------------------------------
// A block of four vertices is 96 bytes = six 16-byte chunks; there is no packed fp64 multiply, so the scan is scalar
// DMUL/DADD.  Same structure as the fp32 scan: one block ahead, (max, block) tracking, exact recovery.
struct Block4d {
  double2 a, b, c, d, e, f;  // x0 y0 | z0 x1 | y1 z1 | x2 y2 | z2 x3 | y3 z3
};
OGJK_D Block4d load_block(const double2* chunk, int g) {
  Block4d k;
  k.a = chunk[6 * g];
  k.b = chunk[6 * g + 1];
  k.c = chunk[6 * g + 2];
  k.d = chunk[6 * g + 3];
  k.e = chunk[6 * g + 4];
  k.f = chunk[6 * g + 5];
  return k;
}
OGJK_D void dots4(const Block4d& k, const V3<double>& D, double (&dd)[4]) {
  dd[0] = add_rn(add_rn(mul_rn(k.a.x, D.x), mul_rn(k.a.y, D.y)), mul_rn(k.b.x, D.z));
  dd[1] = add_rn(add_rn(mul_rn(k.b.y, D.x), mul_rn(k.c.x, D.y)), mul_rn(k.c.y, D.z));
  dd[2] = add_rn(add_rn(mul_rn(k.d.x, D.x), mul_rn(k.d.y, D.y)), mul_rn(k.e.x, D.z));
  dd[3] = add_rn(add_rn(mul_rn(k.e.y, D.x), mul_rn(k.f.x, D.y)), mul_rn(k.f.y, D.z));
}
OGJK_D double max4(const double (&dd)[4]) { return fmax(fmax(dd[0], dd[1]), fmax(dd[2], dd[3])); }
OGJK_D void support_slot(const double* body, int nv, const V3<double>& d, unsigned, V3<double>& sup, int& sup_idx) {
  const double2* chunk = reinterpret_cast<const double2*>(body);
  double best = -INFINITY;
  int bg = 0;
  const int groups = nv >> 2;
  Block4d k0 = load_block(chunk, 0);
#pragma unroll 2
  for (int g = 0; g < groups; ++g) {
    const Block4d n0 = load_block(chunk, g + 1);  // look-ahead; past the end it is discarded
    double da[4];
    dots4(k0, d, da);
    const double ma = max4(da);
    if (ma > best) {  // strict: the earliest block holding the maximum wins
      best = ma;
      bg = g;
    }
    k0 = n0;
  }
  if (best > dot(sup, d)) {
    double dd[4];
    dots4(load_block(chunk, bg), d, dd);
    int k = 3;
    if (dd[2] == best) k = 2;
    if (dd[1] == best) k = 1;
    if (dd[0] == best) k = 0;
    const int idx = 4 * bg + k;
    sup = mk<double>(body[3 * idx], body[3 * idx + 1], body[3 * idx + 2]);
    sup_idx = idx;
  }
}
OGJK_D void support_slots_both(const double* b1, const double* b2, int nv, const V3<double>& v, unsigned zero,
                               V3<double>& sup1, int& idx1, V3<double>& sup2, int& idx2) {
  support_slot(b1, nv, vneg(v), zero, sup1, idx1);
  support_slot(b2, nv, v, zero, sup2, idx2);
}

// integers travel through the finisher's records as bit patterns of T
OGJK_D float rec_from_int(float, unsigned v) { return __uint_as_float(v); }
OGJK_D double rec_from_int(double, unsigned v) { return __longlong_as_double((long long)v); }
OGJK_D unsigned rec_to_uint(float x) { return __float_as_uint(x); }
OGJK_D unsigned rec_to_uint(double x) { return (unsigned)__double_as_longlong(x); }

template <typename T>
struct SlotFetch {
  const T* b1;
  const T* b2;
  OGJK_D V3<T> operator()(int body, int i) const {
    const T* c = body ? b2 : b1;
    return mk<T>(c[3 * i], c[3 * i + 1], c[3 * i + 2]);
  }
};

template <typename T>
struct SlotFetchPacked {  // SoA-4 packed slots (fp32 only)
  const T* b1;
  const T* b2;
  OGJK_D V3<T> operator()(int body, int i) const {
    const T* c = (body ? b2 : b1) + 12 * (i >> 2) + (i & 3);
    return mk<T>(c[0], c[4], c[8]);
  }
};

constexpr int kSlotThreads = 128;
constexpr uint32_t kSlotTableBytes = (kUnifiedSize * 2u + 15u) & ~15u;
constexpr uint32_t kSlotFixedBytes = kSlotThreads * 8u + kSlotTableBytes;  // mbarriers + table
constexpr uint32_t kSlotPadBytes = 192;  // look-ahead loads of the last slot stay inside the allocation

// bytes of one slot: both vertex sets, rounded so that (bytes / 16) is odd
__host__ __device__ inline uint32_t slot_bytes(int nv1, int nv2, int esize = 4) {
  uint32_t units = (uint32_t)(nv1 + nv2) * 3u * (uint32_t)esize / 16u;
  if ((units & 1u) == 0) units += 1;
  return units * 16u;
}

// pool [count][nv][3] -> [count][nv / 4][3][4]; one thread per vertex
__global__ void pack_pool_kernel(const float* __restrict__ in, float* __restrict__ out, int nv, long long total_vertices) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= total_vertices) return;
  const long long body = v / nv;
  const int i = (int)(v - body * nv);
  const float* src = in + 3 * v;
  float* dst = out + body * nv * 3 + 12 * (i >> 2) + (i & 3);
  dst[0] = src[0];
  dst[4] = src[1];
  dst[8] = src[2];
}

// EQ: both bodies have the same vertex count (selects the interleaved two-body scan; the other scan is not even
// instantiated then, which keeps the loop body small for the instruction cache)
// PK: the bodies are SoA-4 packed (x0..x3 | y0..y3 | z0..z3 per four vertices; fp32, equal vertex counts): the pool of an
// indexed batch re-packed on the device by pack_pool_kernel
template <typename T, bool EQ, bool PK = false>
__global__ void __launch_bounds__(kSlotThreads)
gjk_slots_kernel(const T* __restrict__ coord1, const T* __restrict__ coord2, int nv1, int nv2,
                 SimplexT<T>* __restrict__ simplices, T* __restrict__ distances, unsigned n,
                 const uint16_t* __restrict__ utab_g, unsigned* __restrict__ ticket, unsigned zero,
                 const CollisionPair* __restrict__ pairs) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t sbytes = slot_bytes(nv1, nv2, (int)sizeof(T));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // one mbarrier per thread
  uint16_t* utab = reinterpret_cast<uint16_t*>(smem_raw + kSlotThreads * 8u);
  unsigned char* slots = smem_raw + kSlotFixedBytes;
  const int tid = threadIdx.x, lane = tid & 31;
  const T* s1 = reinterpret_cast<const T*>(slots + (size_t)tid * sbytes);
  const T* s2 = s1 + 3 * nv1;
  const uint32_t bar = smem_addr(&bars[tid]);
  const uint32_t dst1 = smem_addr(s1), dst2 = smem_addr(s2);
  const uint32_t bytes1 = (uint32_t)nv1 * 3u * (uint32_t)sizeof(T), bytes2 = (uint32_t)nv2 * 3u * (uint32_t)sizeof(T);

  for (int i = tid; i < kUnifiedSize / 2; i += kSlotThreads)
    reinterpret_cast<uint32_t*>(utab)[i] = __ldg(reinterpret_cast<const uint32_t*>(utab_g) + i);
  mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_proxy_async();
  __syncthreads();

  enum { kNeedWork = 0, kLoading = 1, kRunning = 2, kDone = 3 };
  int state = kNeedWork;
  uint32_t parity = 0;
  unsigned pair = 0;
  TicketFeed feed;
  feed.init(ticket, pairs, n, lane);
  GjkState<T> g;

  for (;;) {
    // ---- hand out work (two passes: a request may straddle two ticket chunks) --------------------------------------
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      unsigned t = 0;
      int i1 = 0, i2 = 0;
      if (feed.take(state == kNeedWork, ticket, pairs, n, lane, t, i1, i2)) {
        if (t < n) {
          pair = t;
          fence_proxy_async();  // this thread's earlier generic-proxy reads of the slot precede the async writes
          mbar_arrive_expect_tx(bar, bytes1 + bytes2);
          tma_bulk_load(dst1, coord1 + (size_t)i1 * nv1 * 3, bytes1, bar);
          tma_bulk_load(dst2, coord2 + (size_t)i2 * nv2 * 3, bytes2, bar);
          state = kLoading;
        } else {
          state = kDone;
        }
      }
    }
    if (__all_sync(0xffffffffu, state == kDone)) break;

    if (state == kLoading && mbar_test_wait(bar, parity)) {
      parity ^= 1u;
      if (PK) gjk_init(g, mk<T>(s1[0], s1[4], s1[8]), mk<T>(s2[0], s2[4], s2[8]));
      else gjk_init(g, mk<T>(s1[0], s1[1], s1[2]), mk<T>(s2[0], s2[1], s2[2]));
      state = kRunning;
    }
    if (state == kRunning) {
      ++g.k;
      if constexpr (PK) {
        support_slots_both_pk(s1, s2, nv1, g.v, zero, g.sup1, g.idx1, g.sup2, g.idx2);
      } else if (EQ) {
        support_slots_both(s1, s2, nv1, g.v, zero, g.sup1, g.idx1, g.sup2, g.idx2);
      } else {
        support_slot(s1, nv1, vneg(g.v), zero, g.sup1, g.idx1);
        support_slot(s2, nv2, g.v, zero, g.sup2, g.idx2);
      }
      if (gjk_advance_u(g, utab)) {
        V3<T> w1, w2;
        if constexpr (PK) {
          SlotFetchPacked<T> fetch{s1, s2};
          gjk_witnesses(fetch, g.S, w1, w2);
        } else {
          SlotFetch<T> fetch{s1, s2};
          gjk_witnesses(fetch, g.S, w1, w2);
        }
        store_result(simplices + pair, distances + pair, g, w1, w2);
        state = kNeedWork;
      }
    }
  }
}

// Slot layout of the warp-specialised kernel.  One lane per pair: body 2 follows body 1, odd 16-byte stride (as above).
// Two lanes per pair: lanes 2p / 2p+1 read body 1 / body 2 of slot p at the same time, so a quarter warp touches four
// slots x two bodies; with body 2 at an ODD 16-byte offset and a stride that is 2 (mod 4) in 16-byte units the eight
// 128-bit loads fall into eight different bank groups.
struct SlotLayout {
  uint32_t stride, body2;  // bytes
};
__host__ __device__ inline SlotLayout ws_slot_layout(int nv1, int nv2, int lp, int esize = 4) {
  SlotLayout L;
  const uint32_t vb = 3u * (uint32_t)esize;  // bytes per vertex
  if (lp == 1) {
    L.stride = slot_bytes(nv1, nv2, esize);
    L.body2 = (uint32_t)nv1 * vb;
  } else {
    uint32_t u1 = ((uint32_t)nv1 * vb + 15u) / 16u;
    if ((u1 & 1u) == 0) u1 += 1;
    uint32_t u = u1 + ((uint32_t)nv2 * vb + 15u) / 16u;
    while ((u & 3u) != 2u) ++u;
    L.stride = u * 16u;
    L.body2 = u1 * 16u;
  }
  return L;
}

// =====================================================================================================================
// Warp-specialised variant (large slots: one warp per scheduler).
//
// profiles/r1d_gjk_slots_v2.txt: with 64+64 vertices only 4 warps fit an SM and the self-service kernel above spends
// 24 % of its time issuing TMA copies (ptxas serialises cp.async.bulk over the lanes: its operands are uniform
// registers), 5 % polling/initialising and 20 % in the witness + store code at 2 active lanes -- all on the critical
// path of the only warp its scheduler has.  Here those jobs move to two helper warps that run in the issue slots the
// compute warps leave idle (issue utilisation is ~30 %):
//   * compute warps (CW x 32 threads, one slot each): wait for the slot's mbarrier, iterate (scans + lane-uniform
//     step), and when the pair terminates copy what the witness stage needs (simplex, v, the <= 8 source vertices)
//     into a record of a shared-memory ring and flag the slot FREE;
//   * the loader warp polls the slot flags, draws tickets (chunked global atomic) and issues the TMA refills;
//   * the finisher warp consumes ring records a warp at a time, one record per lane: witnesses, gkSimplex + distance
//     stores, and -- for the fused GJK+EPA entry points -- the EPA gate (EPA.c:369-373): separated pairs get their
//     contact normal from the witnesses, colliding ones are appended to the EPA queue.
// Synchronisation: slot flag FREE/BUSY/EXIT (plain shared words, fences), per-slot mbarrier for TMA completion (the
// loader's arrive.expect_tx releases its `pair_of` store to the waiting compute thread), ring with a reservation
// counter (shared atomic), per-record generation flags and a consumer-published head.
// ring capacity: 64 records, 32 when the CTA has only 64 slots (keeps 64+64-vertex fp64 slots within shared memory)
__host__ __device__ constexpr int ring_records(int nslots) { return nslots >= 128 ? 64 : 32; }
constexpr int kRecWords = 49;  // odd stride: lanes writing/reading consecutive records hit distinct banks
// record layout (words): 0 pair | 1 n | 2..4 v | 5+5k.. slot k: p.xyz, i1, i2 | 25+6k.. slot k: body-1 xyz, body-2 xyz
enum : unsigned { kSlotFree = 0u, kSlotBusy = 1u, kSlotExit = 2u };

__host__ __device__ constexpr uint32_t ws_fixed_bytes(int nslots, int esize = 4) {
  // mbarriers | table | ctrl | pair_of | ring control (16 B) | ready flags | ring (kRecWords elements of T per record)
  return (uint32_t)nslots * 8u + kSlotTableBytes + (uint32_t)nslots * 4u * 2u + 16u + ring_records(nslots) * 4u +
         (((uint32_t)ring_records(nslots) * kRecWords * (uint32_t)esize + 15u) & ~15u);
}

OGJK_D unsigned ld_vol(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }
OGJK_D void st_vol(unsigned* p, unsigned v) { *reinterpret_cast<volatile unsigned*>(p) = v; }

template <typename T>
struct RecordFetch {  // vertex "index" = original simplex slot in bits 30..31 (see the finisher)
  const T* verts;  // record words 25..48
  OGJK_D V3<T> operator()(int body, int i) const {
    const T* c = verts + 6 * ((unsigned)i >> 30) + 3 * body;
    return mk<T>(c[0], c[1], c[2]);
  }
};

// CW compute warps; LP lanes per pair.  LP = 1: a thread per pair.  LP = 2 (slots so large that only 128 fit an SM):
// the two lanes of a pair each scan ONE body and exchange the support points with a shuffle, everything else is
// evaluated redundantly on both -- twice the warps for the same shared memory, so each scheduler has a second warp to
// issue from while the first waits on a dependency.
// IDX: pairs are gkCollisionPair records into one pool (ticket feed with record prefetch); otherwise pair t is
// (coord1[t], coord2[t]) and the loader draws plain ticket ranges.
template <typename T, int CW, int LP, bool EQ, bool IDX>
__global__ void __launch_bounds__((CW + 2) * 32)
gjk_slots_ws_kernel(const T* __restrict__ coord1, const T* __restrict__ coord2, int nv1, int nv2,
                    SimplexT<T>* __restrict__ simplices, T* __restrict__ distances, unsigned n,
                    const uint16_t* __restrict__ utab_g, unsigned* __restrict__ ticket, unsigned zero,
                    T* __restrict__ normals, int* __restrict__ epa_queue, int* __restrict__ epa_count,
                    const CollisionPair* __restrict__ pairs, unsigned dense_chunk) {
  constexpr int kCompute = CW * 32;
  constexpr int kSlots = kCompute / LP;
  constexpr int kRingRecords = ring_records(kSlots);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const SlotLayout lay = ws_slot_layout(nv1, nv2, LP, (int)sizeof(T));
  const uint32_t sbytes = lay.stride;
  unsigned char* sp = smem_raw;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sp);
  sp += kSlots * 8;
  uint16_t* utab = reinterpret_cast<uint16_t*>(sp);
  sp += kSlotTableBytes;
  unsigned* ctrl = reinterpret_cast<unsigned*>(sp);
  sp += kSlots * 4;
  unsigned* pair_of = reinterpret_cast<unsigned*>(sp);
  sp += kSlots * 4;
  unsigned* ring_ctl = reinterpret_cast<unsigned*>(sp);  // [0] tail (reserved), [1] head (consumed), [2] exited warps
  sp += 16;
  unsigned* ready = reinterpret_cast<unsigned*>(sp);
  sp += kRingRecords * 4;
  T* ring = reinterpret_cast<T*>(sp);
  unsigned char* slots = smem_raw + ws_fixed_bytes(kSlots, (int)sizeof(T));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bytes1 = (uint32_t)nv1 * 3u * (uint32_t)sizeof(T), bytes2 = (uint32_t)nv2 * 3u * (uint32_t)sizeof(T);

  for (int i = tid; i < kUnifiedSize / 2; i += (CW + 2) * 32)
    reinterpret_cast<uint32_t*>(utab)[i] = __ldg(reinterpret_cast<const uint32_t*>(utab_g) + i);
  if (tid < kSlots) {
    mbar_init(smem_addr(&bars[tid]), 1);
    ctrl[tid] = kSlotFree;
    pair_of[tid] = 0;
  }
  for (int i = tid; i < kRingRecords; i += (CW + 2) * 32) ready[i] = 0;
  if (tid < 4) ring_ctl[tid] = 0;
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_proxy_async();
  __syncthreads();

  if (warp < CW) {
    // ================================================ compute ================================================
    const int cslot = tid / LP, half = tid % LP;
    const T* s1 = reinterpret_cast<const T*>(slots + (size_t)cslot * sbytes);
    const T* s2 = reinterpret_cast<const T*>(slots + (size_t)cslot * sbytes + lay.body2);
    const uint32_t bar = smem_addr(&bars[cslot]);
    enum { kWait = 0, kRun = 1, kExit = 2 };
    int state = kWait;
    uint32_t parity = 0;
    unsigned pair = 0;
    GjkState<T> g;
    // The loop is ROTATED: a trip is [sub-algorithm step of the iteration begun last trip] -> [take a new pair if the
    // slot has been refilled] -> [support scans + the two exit pre-tests of the next iteration] -> [retire].  Most
    // pairs end at the pre-tests (every separated pair does), so their slot is handed back right after the scans and
    // the refill overlaps the other lanes' sub-algorithm step.  The arithmetic per pair is the same sequence as in
    // gjk_advance_u.  (Measured: 0.769 against 0.774 ms on config 2 -- the refill, ~2 us from HBM, is still longer
    // than the sub-algorithm step, profiles/r1e_experiments.txt.)
    bool need_sub = false;  // this lane passed the pre-tests last trip and owes the sub-algorithm step
    for (;;) {
      bool fin_sub = false;
      if (state == kRun && need_sub) fin_sub = gjk_substep_u(g, utab);
      need_sub = false;
      if (state == kWait) {
        if (mbar_test_wait(bar, parity)) {
          parity ^= 1u;
          pair = ld_vol(&pair_of[cslot]);
          gjk_init(g, mk<T>(s1[0], s1[1], s1[2]), mk<T>(s2[0], s2[1], s2[2]));
          state = kRun;
        } else if (ld_vol(&ctrl[cslot]) == kSlotExit) {
          state = kExit;
        }
      }
      if (LP == 2) {  // both lanes of a pair must agree (they polled the same words, but not atomically)
        const int other = __shfl_xor_sync(0xffffffffu, state, 1);
        if (other != state) {  // one saw the barrier flip (or EXIT), the other did not yet: retry next trip
          if (state == kRun) parity ^= 1u;
          state = kWait;
        }
      }
      if (__all_sync(0xffffffffu, state == kExit)) break;
      if (!__any_sync(0xffffffffu, state == kRun)) __nanosleep(32);  // start-up / drain: nothing loaded yet
      bool finished = fin_sub;
      const unsigned runm = __ballot_sync(0xffffffffu, state == kRun && !fin_sub);  // both lanes of a pair, or neither
      if (state == kRun && !fin_sub) {
        ++g.k;
        if (LP == 1) {
          if (EQ) {
            support_slots_both(s1, s2, nv1, g.v, zero, g.sup1, g.idx1, g.sup2, g.idx2);
          } else {
            support_slot(s1, nv1, vneg(g.v), zero, g.sup1, g.idx1);
            support_slot(s2, nv2, g.v, zero, g.sup2, g.idx2);
          }
        } else {
          // this lane's body: half 0 scans body 1 along -v, half 1 scans body 2 along +v
          const T* body = half ? s2 : s1;
          const int nvb = half ? nv2 : nv1;
          const V3<T> d = half ? g.v : vneg(g.v);
          V3<T> sup = half ? g.sup2 : g.sup1;
          int sidx = half ? g.idx2 : g.idx1;
          support_slot(body, nvb, d, zero, sup, sidx);
          const T ox = __shfl_xor_sync(runm, sup.x, 1), oy = __shfl_xor_sync(runm, sup.y, 1),
                  oz = __shfl_xor_sync(runm, sup.z, 1);
          const int oi = __shfl_xor_sync(runm, sidx, 1);
          const V3<T> osup = mk<T>(ox, oy, oz);
          g.sup1 = half ? osup : sup;
          g.sup2 = half ? sup : osup;
          g.idx1 = half ? oi : sidx;
          g.idx2 = half ? sidx : oi;
        }
        finished = gjk_converged_u(g);
        need_sub = !finished;
      }
      // (Measured and dropped in round 2: SCANNER WARPS -- 2 x CW extra warps, one thread per (slot, body), scanning on
      // request through shared-memory mailboxes while the slot's owner thread only merges, tests and runs the
      // sub-algorithm step; bit-exact, 60 % issue utilisation, but 1.01-1.23 ms against 0.77 ms: the two hand-overs per
      // iteration cost more than the scan they take off the owner's chain -- profiles/r2_experiments.txt, ncu
      // summaries profiles/r2[c-e]_gjk_slots_ws_scanners_*.txt, code in commits d972fc6..3664304.)
      // (Measured and dropped: a second retire call site in front of the sub-algorithm; pulling upcoming pairs into L2
      // -- cp.async.bulk.prefetch.L2 or per-line prefetch.global.L2 from the loader, any distance; staging buffers
      // behind the slots; extra compute warps with register-resident pairs.  profiles/r1c_gjk_kernels_ab.txt,
      // profiles/r1e_experiments.txt.)
      auto retire = [&](bool done) {
        const unsigned fin = __ballot_sync(0xffffffffu, done && half == 0);
        if (!fin) return;
        const unsigned cnt = __popc(fin);
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&ring_ctl[0], cnt);
        base = __shfl_sync(0xffffffffu, base, 0);
        while ((int)(base + cnt - ld_vol(&ring_ctl[1])) > kRingRecords) __nanosleep(64);  // ring full: wait for space
        const unsigned idx = base + __popc(fin & ((1u << (lane & ~(LP - 1))) - 1u));
        if (done) {
          T* rec = ring + (size_t)(idx % kRingRecords) * kRecWords;
          const SV<T>* sv[4] = {&g.S.s0, &g.S.s1, &g.S.s2, &g.S.s3};
          // all slot reads first, then all record writes: both are shared memory, so the compiler keeps their order
          // and a load placed after a store would wait out its full latency before the next store can issue
          T vert[4][6];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (LP == 1 || (k >> 1) == half) {
              const SV<T>& q = *sv[k];
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                vert[k][c] = s1[3 * q.i1 + c];
                vert[k][3 + c] = s2[3 * q.i2 + c];
              }
            }
          }
          if (half == 0) {
            rec[0] = rec_from_int(T(0), pair);
            rec[1] = rec_from_int(T(0), (unsigned)g.S.n);
            rec[2] = g.v.x;
            rec[3] = g.v.y;
            rec[4] = g.v.z;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (LP == 1 || (k >> 1) == half) {
              const SV<T>& q = *sv[k];
              rec[5 + 5 * k + 0] = q.p.x;
              rec[5 + 5 * k + 1] = q.p.y;
              rec[5 + 5 * k + 2] = q.p.z;
              rec[5 + 5 * k + 3] = rec_from_int(T(0), (unsigned)q.i1);
              rec[5 + 5 * k + 4] = rec_from_int(T(0), (unsigned)q.i2);
#pragma unroll
              for (int c = 0; c < 6; ++c) rec[25 + 6 * k + c] = vert[k][c];
            }
          }
          __threadfence_block();  // record + this thread's slot reads before the two flags
        }
        __syncwarp();
        if (done) {
          if (half == 0) {
            st_vol(&ready[idx % kRingRecords], idx / kRingRecords + 1u);
            st_vol(&ctrl[cslot], kSlotFree);
          }
          state = kWait;
        }
      };
      retire(finished);
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      atomicAdd(&ring_ctl[2], 1u);
    }
  } else if (warp == CW) {
    // ================================================ loader =================================================
    TicketFeed feed;
    if (IDX) feed.init(ticket, pairs, n, lane);
    unsigned tk_next = 0, tk_end = 0;  // dense batches: reserved ticket range (warp-uniform), dense_chunk per atomic
    unsigned exited = 0;               // bit j: slot lane + 32 j has been told to exit
    for (;;) {
      bool any = false;
#pragma unroll  // (a rolled loop makes the kernel 11 % smaller but each poll slower: 0.781 against 0.769 ms)
      for (int j = 0; j < kSlots / 32; ++j) {
        const int s = lane + 32 * j;
        bool want = !((exited >> j) & 1u) && ld_vol(&ctrl[s]) == kSlotFree;
        const unsigned wm = __ballot_sync(0xffffffffu, want);
        if (!wm) continue;  // keep the idle polling loop light: it shares a scheduler with a compute warp
        any = true;
        unsigned nb = 0, avail = 0;
        if (!IDX) {
          avail = tk_end - tk_next;
          if (__popc(wm) > avail) {
            if (lane == 0) nb = atomicAdd(ticket, dense_chunk);
            nb = __shfl_sync(0xffffffffu, nb, 0);
          }
        }
#pragma unroll 1
        for (int pass = 0; pass < (IDX ? 2 : 1); ++pass) {
          unsigned t = 0;
          int i1 = 0, i2 = 0;
          bool served;
          if (IDX) {
            served = feed.take(want, ticket, pairs, n, lane, t, i1, i2);
          } else {
            const unsigned r = __popc(wm & ((1u << lane) - 1u));
            t = r < avail ? tk_next + r : nb + (r - avail);
            i1 = i2 = (int)t;
            served = want;
          }
          if (served) {
            want = false;
            if (t < n) {
              __threadfence_block();  // the owner's last reads of the slot happened before it flagged FREE
              pair_of[s] = t;
              st_vol(&ctrl[s], kSlotBusy);
              const uint32_t bar = smem_addr(&bars[s]);
              const uint32_t dst1 = smem_addr(slots + (size_t)s * sbytes), dst2 = dst1 + lay.body2;
              fence_proxy_async();
              mbar_arrive_expect_tx(bar, bytes1 + bytes2);  // release: pair_of is visible to the waiting thread
              tma_bulk_load(dst1, coord1 + (size_t)i1 * nv1 * 3, bytes1, bar);
              tma_bulk_load(dst2, coord2 + (size_t)i2 * nv2 * 3, bytes2, bar);
            } else {
              st_vol(&ctrl[s], kSlotExit);
              exited |= 1u << j;
            }
          }
        }
        if (!IDX) {
          const unsigned cnt = __popc(wm);
          if (cnt > avail) {
            tk_next = nb + (cnt - avail);
            tk_end = nb + dense_chunk;
          } else {
            tk_next += cnt;
          }
        }
      }
      if (__all_sync(0xffffffffu, exited == (1u << (kSlots / 32)) - 1u)) break;
      if (!any) __nanosleep(100);
    }
  } else {
    // ================================================ finisher ===============================================
    unsigned cur = 0;
    for (;;) {
      const unsigned idx = cur + lane;
      const bool rdy = ld_vol(&ready[idx % kRingRecords]) == idx / kRingRecords + 1u;
      const unsigned m = __ballot_sync(0xffffffffu, rdy);
      const int c = (m == 0xffffffffu) ? 32 : (__ffs(~m) - 1);  // records ready in order from `cur`
      if (c == 0) {
        if (ld_vol(&ring_ctl[2]) == (unsigned)CW && ld_vol(&ring_ctl[0]) == cur) break;
        __nanosleep(100);
        continue;
      }
      __threadfence_block();
      if (lane < c) {
        const T* rec = ring + (size_t)(idx % kRingRecords) * kRecWords;
        const unsigned pair = rec_to_uint(rec[0]);
        GjkState<T> g;
        g.S.n = (int)rec_to_uint(rec[1]);
        g.v = mk<T>(rec[2], rec[3], rec[4]);
        SV<T>* sv[4] = {&g.S.s0, &g.S.s1, &g.S.s2, &g.S.s3};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sv[k]->p = mk<T>(rec[5 + 5 * k], rec[6 + 5 * k], rec[7 + 5 * k]);
          // tag (bits 30..31): which record vertex this slot came with -- survives the witness stage's slot shuffles
          sv[k]->i1 = (int)(rec_to_uint(rec[8 + 5 * k]) | ((unsigned)k << 30));
          sv[k]->i2 = (int)(rec_to_uint(rec[9 + 5 * k]) | ((unsigned)k << 30));
        }
        RecordFetch<T> fetch{rec + 25};
        V3<T> w1, w2;
        gjk_witnesses(fetch, g.S, w1, w2);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sv[k]->i1 &= 0x3fffffff;
          sv[k]->i2 &= 0x3fffffff;
        }
        store_result(simplices + pair, distances + pair, g, w1, w2);
        if (normals) {  // fused EPA gate
          const T dist = sqrt_rn(norm2(g.v));
          bool collide = !(dist > Tol<T>::eps());
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

// gjk_slots.cuh -- persistent "slot" GJK kernel for uniform fp32 batches: the Blackwell-native hot path.
//
// ONE THREAD OWNS ONE PAIR for every phase (support scans, exit tests, sub-algorithm, witnesses), so no instruction
// is executed redundantly by cooperating lanes, and the kernel is persistent so no lane idles while others iterate:
//   * each thread has a private shared-memory slot holding its pair's two vertex sets exactly as they lie in HBM
//     (xyz interleaved).  A finished thread takes the next pair index from a global ticket and refills its slot with
//     two TMA bulk copies (cp.async.bulk global->shared, completion on the slot's own mbarrier); while the copy is
//     in flight the other 31 lanes keep iterating.  This is the work queue that rebalances pairs whose iteration
//     counts diverge (1..25 iterations, mean 3.8 at 64 vertices).  Tickets are drawn per warp in chunks of 64 (one
//     atomic per ~8 warp iterations instead of one per iteration), and the chunk that will be needed
//     `prefetch_ahead` pairs later is pulled into L2 with cp.async.bulk.prefetch so that refills hit L2, not HBM;
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
OGJK_D void tma_prefetch_l2(const void* src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
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

struct SlotFetch {
  const float* b1;
  const float* b2;
  OGJK_D V3<float> operator()(int body, int i) const {
    const float* c = body ? b2 : b1;
    return mk<float>(c[3 * i], c[3 * i + 1], c[3 * i + 2]);
  }
};

constexpr int kSlotThreads = 128;
constexpr int kTicketChunk = 64;
constexpr uint32_t kSlotTableBytes = (kUnifiedSize * 2u + 15u) & ~15u;
constexpr uint32_t kSlotFixedBytes = kSlotThreads * 8u + kSlotTableBytes;  // mbarriers + table
constexpr uint32_t kSlotPadBytes = 96;  // look-ahead loads of the last slot stay inside the allocation

// bytes of one slot: both vertex sets, rounded so that (bytes / 16) is odd
__host__ __device__ inline uint32_t slot_bytes(int nv1, int nv2) {
  uint32_t units = (uint32_t)(nv1 + nv2) * 12u / 16u;
  if ((units & 1u) == 0) units += 1;
  return units * 16u;
}

__global__ void __launch_bounds__(kSlotThreads)
gjk_slots_kernel(const float* __restrict__ coord1, const float* __restrict__ coord2, int nv1, int nv2,
                 SimplexT<float>* __restrict__ simplices, float* __restrict__ distances, unsigned n,
                 const uint16_t* __restrict__ utab_g, unsigned* __restrict__ ticket, unsigned prefetch_ahead,
                 unsigned zero) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t sbytes = slot_bytes(nv1, nv2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // one mbarrier per thread
  uint16_t* utab = reinterpret_cast<uint16_t*>(smem_raw + kSlotThreads * 8u);
  unsigned char* slots = smem_raw + kSlotFixedBytes;
  const int tid = threadIdx.x, lane = tid & 31;
  const float* s1 = reinterpret_cast<const float*>(slots + (size_t)tid * sbytes);
  const float* s2 = s1 + 3 * nv1;
  const uint32_t bar = smem_addr(&bars[tid]);
  const uint32_t dst1 = smem_addr(s1), dst2 = smem_addr(s2);
  const uint32_t bytes1 = (uint32_t)nv1 * 12u, bytes2 = (uint32_t)nv2 * 12u;

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
  unsigned tk_next = 0, tk_end = 0;  // this warp's reserved ticket range (warp-uniform)
  GjkState<float> g;

  for (;;) {
    // ---- hand out work: per-warp ticket chunks ---------------------------------------------------------------
    const unsigned want = __ballot_sync(0xffffffffu, state == kNeedWork);
    if (want) {
      const unsigned cnt = __popc(want), avail = tk_end - tk_next;
      unsigned nb = 0;
      if (cnt > avail) {
        if (lane == 0) nb = atomicAdd(ticket, (unsigned)kTicketChunk);
        nb = __shfl_sync(0xffffffffu, nb, 0);
        if (prefetch_ahead) {  // the chunk `prefetch_ahead` pairs further on: two pairs per lane, contiguous in HBM
          const unsigned long long p = (unsigned long long)nb + prefetch_ahead + 2u * lane;
          if (p + 2 <= n) {
            tma_prefetch_l2(coord1 + p * nv1 * 3, 2u * bytes1);
            tma_prefetch_l2(coord2 + p * nv2 * 3, 2u * bytes2);
          }
        }
      }
      if (state == kNeedWork) {
        const unsigned r = __popc(want & ((1u << lane) - 1u));
        const unsigned t = r < avail ? tk_next + r : nb + (r - avail);
        if (t < n) {
          pair = t;
          fence_proxy_async();  // this thread's earlier generic-proxy reads of the slot precede the async writes
          mbar_arrive_expect_tx(bar, bytes1 + bytes2);
          tma_bulk_load(dst1, coord1 + (size_t)t * nv1 * 3, bytes1, bar);
          tma_bulk_load(dst2, coord2 + (size_t)t * nv2 * 3, bytes2, bar);
          state = kLoading;
        } else {
          state = kDone;
        }
      }
      if (cnt > avail) {
        tk_next = nb + (cnt - avail);
        tk_end = nb + kTicketChunk;
      } else {
        tk_next += cnt;
      }
    }
    if (__all_sync(0xffffffffu, state == kDone)) break;

    if (state == kLoading && mbar_test_wait(bar, parity)) {
      parity ^= 1u;
      gjk_init(g, mk<float>(s1[0], s1[1], s1[2]), mk<float>(s2[0], s2[1], s2[2]));
      state = kRunning;
    }
    if (state == kRunning) {
      ++g.k;
      support_slot(s1, nv1, vneg(g.v), zero, g.sup1, g.idx1);
      support_slot(s2, nv2, g.v, zero, g.sup2, g.idx2);
      if (gjk_advance_u(g, utab)) {
        SlotFetch fetch{s1, s2};
        V3<float> w1, w2;
        gjk_witnesses(fetch, g.S, w1, w2);
        store_result(simplices + pair, distances + pair, g, w1, w2);
        state = kNeedWork;
      }
    }
  }
}

}  // namespace ogjk

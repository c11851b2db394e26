// gjk_slots.cuh -- persistent "slot" GJK kernel for uniform fp32 batches: the Blackwell-native hot path.
//
// Why a third kernel: profiles/r1_gjk_uniform.txt shows that once the support scan is cheap, the scalar part of a
// GJK iteration (exit tests + signed-volumes sub-algorithm) dominates, and in the L-lanes-per-pair kernels that
// part is executed redundantly by all L lanes while finished pairs keep their lanes idle (16 of 32 lanes active on
// average).  Here ONE THREAD OWNS ONE PAIR for every phase, so no instruction is redundant, and the kernel is
// persistent so no lane idles:
//   * each thread has a private shared-memory slot holding its pair's two vertex sets exactly as they lie in HBM
//     (xyz interleaved).  A finished thread takes the next pair index from a global ticket (warp-aggregated atomic)
//     and refills its slot with two TMA bulk copies (cp.async.bulk global->shared, completion on the slot's own
//     mbarrier); while the copy is in flight the other 31 lanes keep iterating.  This is the work queue that
//     rebalances pairs whose iteration counts diverge (1..25 iterations, mean 3.8 at 64 vertices).
//   * slot stride is an odd multiple of 16 bytes, so the 128-bit shared loads of the 32 lanes of a warp (each in
//     its own slot) are bank-conflict free;
//   * the support scan walks the slot four vertices (three 128-bit loads) at a time with packed FMUL2 products and
//     scalar adds (see gjk_uniform.cuh for why the adds are not packed), keeps only the running maximum and the
//     index of the winning 4-vertex block (FMNMX3 + one compare/select pair per block), and recovers the exact
//     lowest winning index from that one block afterwards (SURVEY.md Appendix A.2 tie-break);
//   * exit tests, table-driven sub-algorithm and witnesses are the shared per-thread core (gjk_core.cuh); witness
//     vertices are fetched from the slot, so global memory is touched once per vertex.
// HBM traffic is the algorithmic minimum: every vertex byte is read once (by TMA), every result byte written once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gjk_core.cuh"
#include "gjk_generic.cuh"
#include "gjk_uniform.cuh"
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

// dots of the four vertices held in three consecutive 16-byte chunks (x0 y0 z0 x1 | y1 z1 x2 y2 | z2 x3 y3 z3)
OGJK_D void dots4(const float4& A, const float4& B, const float4& C, const V3<float>& d, float (&out)[4]) {
  const float v[12] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, C.x, C.y, C.z, C.w};
  float p[12];
  products(v, d, p);
#pragma unroll
  for (int j = 0; j < 4; ++j) out[j] = add_rn(add_rn(p[3 * j], p[3 * j + 1]), p[3 * j + 2]);
}

// Support search of one body by ONE thread over its shared-memory slot.  `body` points at nv*3 floats, nv % 4 == 0.
OGJK_D void support_slot(const float* body, int nv, const V3<float>& d, V3<float>& sup, int& sup_idx) {
  const float4* chunk = reinterpret_cast<const float4*>(body);
  float best = -INFINITY;
  int bg = 0;
  const int groups = nv >> 2;
#pragma unroll 4
  for (int g = 0; g < groups; ++g) {
    float dd[4];
    dots4(chunk[3 * g], chunk[3 * g + 1], chunk[3 * g + 2], d, dd);
    const float m = fmaxf(fmaxf(dd[0], dd[1]), fmaxf(dd[2], dd[3]));
    if (m > best) {  // strict: the earliest block holding the maximum wins
      best = m;
      bg = g;
    }
  }
  if (best > dot(sup, d)) {
    float dd[4];
    dots4(chunk[3 * bg], chunk[3 * bg + 1], chunk[3 * bg + 2], d, dd);
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

// bytes of one slot: both vertex sets, rounded so that (bytes / 16) is odd
__host__ __device__ inline uint32_t slot_bytes(int nv1, int nv2) {
  uint32_t units = (uint32_t)(nv1 + nv2) * 12u / 16u;
  if ((units & 1u) == 0) units += 1;
  return units * 16u;
}

__global__ void __launch_bounds__(kSlotThreads)
gjk_slots_kernel(const float* __restrict__ coord1, const float* __restrict__ coord2, int nv1, int nv2,
                 SimplexT<float>* __restrict__ simplices, float* __restrict__ distances, int n,
                 const uint32_t* __restrict__ tabs, int* __restrict__ ticket) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t sbytes = slot_bytes(nv1, nv2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // one mbarrier per thread
  unsigned char* slots = smem_raw + kSlotThreads * sizeof(uint64_t);
  const int tid = threadIdx.x, lane = tid & 31;
  const float* s1 = reinterpret_cast<const float*>(slots + (size_t)tid * sbytes);
  const float* s2 = s1 + 3 * nv1;
  const uint32_t bar = smem_addr(&bars[tid]);
  const uint32_t dst1 = smem_addr(s1), dst2 = smem_addr(s2);
  const uint32_t bytes1 = (uint32_t)nv1 * 12u, bytes2 = (uint32_t)nv2 * 12u;
  const uint32_t* t3 = tabs;
  const uint32_t* t2 = tabs + 4096;

  mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_proxy_async();
  __syncwarp();

  enum { kNeedWork = 0, kLoading = 1, kRunning = 2, kDone = 3 };
  int state = kNeedWork;
  uint32_t parity = 0;
  long long pair = -1;
  GjkState<float> g;

  for (;;) {
    // ---- hand out work: warp-aggregated ticket ---------------------------------------------------------------
    const unsigned want = __ballot_sync(0xffffffffu, state == kNeedWork);
    if (want) {
      int base = 0;
      const int leader = __ffs(want) - 1;
      if (lane == leader) base = atomicAdd(ticket, __popc(want));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (state == kNeedWork) {
        const long long t = (long long)base + __popc(want & ((1u << lane) - 1u));
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
    }
    if (__all_sync(0xffffffffu, state == kDone)) break;

    if (state == kLoading && mbar_test_wait(bar, parity)) {
      parity ^= 1u;
      gjk_init(g, mk<float>(s1[0], s1[1], s1[2]), mk<float>(s2[0], s2[1], s2[2]));
      state = kRunning;
    }
    if (state == kRunning) {
      ++g.k;
      support_slot(s1, nv1, vneg(g.v), g.sup1, g.idx1);
      support_slot(s2, nv2, g.v, g.sup2, g.idx2);
      if (gjk_advance(g, t2, t3)) {
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

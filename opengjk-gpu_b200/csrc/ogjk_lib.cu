// ogjk_lib.cu -- the C ABI (include/opengjk_b200.h) over the sm_100a kernels.
//
// Host side of the drop-in boundary: the same allocate / flatten / launch / copy-back / free steps as the
// reference's wrappers (GJK/gpu/openGJK.cu:2787-3311), with checked CUDA calls, 64-bit size arithmetic (the
// reference accumulates coordinate counts in `int`, openGJK.cu:2910-2915), pinned staging and stream-ordered
// copies.  There is deliberately no CPU fallback: without a usable device every compute call fails.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/opengjk_b200.h"
#include "epa_kernel.cuh"
#include "epa_group.cuh"
#include "broadphase.cuh"
#include "contact.cuh"
#include "transform.cuh"
#include "gjk_generic.cuh"
#include "gjk_tables.h"
#include "gjk_slots.cuh"
#include "gjk_slots16.cuh"
#include "gjk_uniform.cuh"
#include "ogjk_types.h"

using namespace ogjk;

namespace {

thread_local std::string t_err;
thread_local cudaStream_t t_stream = nullptr;
thread_local bool t_sync = true;
thread_local long long t_launches = 0;
thread_local const char* t_last_kernel = "";  // name of the last kernel launched through finish_launch

// optional per-stage device timing of the fused GJK+EPA entry points (ogjk_set_timing / ogjk_stage_times): three
// events per call on the launching stream -- before GJK, between the stages, after EPA
struct StageEvents {
  cudaEvent_t e[3];
};
thread_local bool t_timing = false;
thread_local std::vector<StageEvents> t_stage_pool;  // created lazily, re-used
thread_local size_t t_stage_used = 0;

int fail(const char* what, cudaError_t e) {
  char buf[512];
  snprintf(buf, sizeof(buf), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  t_err = buf;
  return (int)e ? (int)e : -1;
}
int fail_msg(const char* msg) {
  t_err = msg;
  return -1;
}

#define OGJK_CK(call)                                 \
  do {                                                \
    cudaError_t e__ = (call);                         \
    if (e__ != cudaSuccess) return fail(#call, e__);  \
  } while (0)

// ---- decision tables, one copy per device ------------------------------------------------------------
constexpr int kMaxDevices = 64;
std::mutex g_tab_mutex;
const uint32_t* g_tabs[kMaxDevices] = {};

int device_tables(const uint32_t** out) {
  int dev = 0;
  OGJK_CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail_msg("device ordinal out of range");
  std::lock_guard<std::mutex> lock(g_tab_mutex);
  if (!g_tabs[dev]) {
    static LeafTables host_tabs;
    static bool built = false;
    if (!built) {
      build_leaf_tables(host_tabs);
      built = true;
    }
    uint32_t* d = nullptr;
    OGJK_CK(cudaMalloc(&d, sizeof(LeafTables)));
    OGJK_CK(cudaMemcpy(d, &host_tabs, sizeof(LeafTables), cudaMemcpyHostToDevice));
    g_tabs[dev] = d;
  }
  *out = g_tabs[dev];
  return 0;
}

// the 16-bit unified table of the lane-uniform iteration (gjk_tables.h), one copy per device
const uint16_t* g_utab[kMaxDevices] = {};
int device_unified_table(const uint16_t** out) {
  int dev = 0;
  OGJK_CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail_msg("device ordinal out of range");
  std::lock_guard<std::mutex> lock(g_tab_mutex);
  if (!g_utab[dev]) {
    static uint16_t host_tab[kUnifiedSize];
    static bool built = false;
    if (!built) {
      build_unified_table(host_tab);
      built = true;
    }
    uint16_t* d = nullptr;
    OGJK_CK(cudaMalloc(&d, sizeof(host_tab)));
    OGJK_CK(cudaMemcpy(d, host_tab, sizeof(host_tab), cudaMemcpyHostToDevice));
    g_utab[dev] = d;
  }
  *out = g_utab[dev];
  return 0;
}

int finish_launch(const char* what) {
  ++t_launches;
  t_last_kernel = what;
  OGJK_CK(cudaGetLastError());
  if (t_sync) {
    cudaError_t e = cudaStreamSynchronize(t_stream);
    if (e != cudaSuccess) return fail(what, e);
  }
  return 0;
}

// Launch configuration of the persistent kernels, memoised per (thread, device, kernel, shared-memory size): the
// dynamic shared-memory opt-in and the occupancy query cost ~10 us of host time per call, which a small batch would
// see as GPU idle time in front of every launch.
struct LaunchCfg {
  int dev;
  const void* fn;
  size_t smem;
  int grid_per_device;  // SMs x resident CTAs per SM
};
thread_local std::vector<LaunchCfg> t_cfgs;
struct SmemAttr {
  int dev;
  const void* fn;
  size_t bytes;  // dynamic shared-memory opt-in currently set for this kernel on this device
};
std::mutex g_attr_mutex;
std::vector<SmemAttr> g_attrs;
template <typename Kernel>
int persistent_grid(Kernel kern, int threads, size_t smem, long long* grid) {
  int dev = 0;
  OGJK_CK(cudaGetDevice(&dev));
  const void* fn = reinterpret_cast<const void*>(kern);
  for (const LaunchCfg& c : t_cfgs)
    if (c.dev == dev && c.fn == fn && c.smem == smem) {
      *grid = c.grid_per_device;
      return 0;
    }
  int sms = 0, per_sm = 0;
  // the opt-in limit is one attribute per kernel: raise it, never lower it -- another shape of the same kernel may
  // already be cached here with a larger request and would otherwise fail to launch later
  if (smem > 48u * 1024u) {
    std::lock_guard<std::mutex> lock(g_attr_mutex);  // process-wide: the attribute belongs to (device, kernel)
    SmemAttr* slot = nullptr;
    for (SmemAttr& a : g_attrs)
      if (a.dev == dev && a.fn == fn) slot = &a;
    if (!slot) {
      g_attrs.push_back(SmemAttr{dev, fn, 0});
      slot = &g_attrs.back();
    }
    if (smem > slot->bytes) {
      OGJK_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      slot->bytes = smem;
    }
  }
  OGJK_CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  OGJK_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  if (per_sm < 1) return fail_msg("persistent kernel does not fit on this device");
  t_cfgs.push_back(LaunchCfg{dev, fn, smem, sms * per_sm});
  *grid = (long long)sms * per_sm;
  return 0;
}

// lanes per pair for the general kernel, from a typical vertex count
int lanes_for(int nv) {
  if (nv <= 16) return 4;
  if (nv <= 96) return 8;
  if (nv <= 384) return 16;
  return 32;
}

template <typename T, typename Source>
int launch_gjk_generic(const Source& src, int n, int nv_hint, SimplexT<T>* d_simplices, T* d_distances) {
  const uint32_t* tabs = nullptr;
  if (int rc = device_tables(&tabs)) return rc;
  const int L = lanes_for(nv_hint);
  const int block = 256;
  const long long threads = (long long)n * L;
  const unsigned grid = (unsigned)((threads + block - 1) / block);
  switch (L) {
    case 4: gjk_generic_kernel<T, 4, Source><<<grid, block, 0, t_stream>>>(src, d_simplices, d_distances, n, tabs); break;
    case 8: gjk_generic_kernel<T, 8, Source><<<grid, block, 0, t_stream>>>(src, d_simplices, d_distances, n, tabs); break;
    case 16: gjk_generic_kernel<T, 16, Source><<<grid, block, 0, t_stream>>>(src, d_simplices, d_distances, n, tabs); break;
    default: gjk_generic_kernel<T, 32, Source><<<grid, block, 0, t_stream>>>(src, d_simplices, d_distances, n, tabs); break;
  }
  return finish_launch("gjk kernel");
}

// Uniform batches (dense [n][V][3] coordinates, V % 4 == 0, 16-byte aligned): register-resident fast kernel.
// Returns 1 if the batch does not qualify (caller falls back to the general kernel), 0 on success, <0 on error.
template <typename T, int L, int VPL>
void launch_uniform_instance(const T* c1, const T* c2, int nv1, int nv2, SimplexT<T>* simp, T* dist, int n,
                             const uint32_t* tabs) {
  const int block = 256;
  const long long threads = (long long)n * L;
  const unsigned grid = (unsigned)((threads + block - 1) / block);
  gjk_uniform_kernel<T, L, VPL><<<grid, block, 0, t_stream>>>(c1, c2, nv1, nv2, simp, dist, n, tabs);
}

// ---- per-(thread, device, stream) scratch ---------------------------------------------------------------------
// The GJK ticket, the EPA counters + queue and the broad-phase / contact-response work buffers are written by the
// kernels of one call and must not be shared by two calls in flight.  Calls on one stream are ordered by the stream;
// calls issued asynchronously (ogjk_set_sync(0)) on DIFFERENT streams of one thread get different buffers because
// the scratch is keyed by (device, stream).  Grow-only; released by ogjk_release_cached_buffers().
struct Scratch {
  int* ptr = nullptr;
  size_t ints = 0;
};
struct StreamScratch {
  int dev = -1;
  cudaStream_t stream = nullptr;
  unsigned* ticket = nullptr;
  Scratch epa, bp, cr, flag, pack;
};
thread_local std::vector<StreamScratch> t_sscratch;
int stream_scratch(StreamScratch** out) {
  int dev = 0;
  OGJK_CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail_msg("device ordinal out of range");
  for (StreamScratch& x : t_sscratch)
    if (x.dev == dev && x.stream == t_stream) {
      *out = &x;
      return 0;
    }
  StreamScratch x;
  x.dev = dev;
  x.stream = t_stream;
  t_sscratch.push_back(x);
  *out = &t_sscratch.back();
  return 0;
}
int scratch_grow(Scratch& sc, size_t ints, int** out) {
  if (sc.ints < ints) {
    if (sc.ptr) cudaFree(sc.ptr);
    sc.ptr = nullptr;
    sc.ints = 0;
    OGJK_CK(cudaMalloc(&sc.ptr, ints * sizeof(int)));
    sc.ints = ints;
  }
  *out = sc.ptr;
  return 0;
}
int ticket_buffer(unsigned** out) {
  StreamScratch* ss = nullptr;
  if (int rc = stream_scratch(&ss)) return rc;
  if (!ss->ticket) OGJK_CK(cudaMalloc(&ss->ticket, sizeof(unsigned)));
  *out = ss->ticket;
  return 0;
}

// ---- persistent slot kernel (fp32, both vertex sets of a pair fit one shared-memory slot) ----------------------

// development override: OGJK_GJK_KERNEL=slots|slotsws|uniform|generic forces one kernel family (A/B measurements)
// (re-read on every call: the parity tests switch it between calls; one getenv costs ~50 ns)
int forced_kernel() {
  const char* e = getenv("OGJK_GJK_KERNEL");
  return !e ? 0 : !strcmp(e, "slots") ? 1 : !strcmp(e, "uniform") ? 2 : !strcmp(e, "generic") ? 3 :
         !strcmp(e, "slotsws") ? 4 : !strcmp(e, "slots16") ? 5 : !strcmp(e, "slotsws32") ? 6 : 0;
}

// Indexed batches over one fp32 pool can have the pool re-packed on the device into SoA-4 blocks (x0..x3 | y0..y3 |
// z0..z3) in front of the self-service slot kernel -- one pass over the pool, which an indexed batch reads hundreds of
// times -- so that the scan runs with packed multiplies AND packed adds (gjk_slots.cuh, dots4_pk: 10 instead of 14
// issue slots per four vertices).  Bit-exact, measured, and NOT the default: config 5 (20 000 x 32 vertices, 3.96 M
// pairs) 1.846 ms packed against 1.803 ms plain (profiles/r2r_ab_pool_pack.txt) -- the kernel waits on dependent
// results, not on issue slots.  OGJK_POOL_PACK=1 turns it on.
bool use_packed_pool(long long, int pool_count) {
  const char* e = getenv("OGJK_POOL_PACK");
  return e && atoi(e) != 0 && pool_count > 0;
}

template <typename T>
int launch_gjk_slots(int n, int nv1, const T* c1, int nv2, const T* c2, SimplexT<T>* simp, T* dist,
                     const CollisionPair* pairs = nullptr, int pool_count = 0) {
  const uint16_t* utab = nullptr;
  if (int rc = device_unified_table(&utab)) return rc;
  unsigned* ticket = nullptr;
  if (int rc = ticket_buffer(&ticket)) return rc;
  const size_t smem = (size_t)kSlotFixedBytes + kSlotPadBytes + (size_t)kSlotThreads * slot_bytes(nv1, nv2, (int)sizeof(T));
  // the interleaved scan needs ~40 more registers: only where shared memory, not registers, bounds occupancy
  const bool eq = sizeof(T) == 4 && nv1 == nv2 && nv1 >= 32;
  auto kern = eq ? gjk_slots_kernel<T, true> : gjk_slots_kernel<T, false>;
  bool packed = false;
  if constexpr (sizeof(T) == 4) {
    if (eq && pairs && c1 == c2 && pool_count > 0 && use_packed_pool(n, pool_count)) {
      StreamScratch* ss = nullptr;
      if (int rc = stream_scratch(&ss)) return rc;
      int* buf = nullptr;
      const long long verts = (long long)pool_count * nv1;
      if (int rc = scratch_grow(ss->pack, (size_t)verts * 3, &buf)) return rc;
      pack_pool_kernel<<<(unsigned)((verts + 255) / 256), 256, 0, t_stream>>>(c1, reinterpret_cast<float*>(buf), nv1, verts);
      ++t_launches;
      OGJK_CK(cudaGetLastError());
      c1 = c2 = reinterpret_cast<const T*>(buf);
      kern = gjk_slots_kernel<T, true, true>;
      packed = true;
    }
  }
  long long grid = 0;
  if (int rc = persistent_grid(kern, kSlotThreads, smem, &grid)) return rc;
  const long long need = ((long long)n + kSlotThreads - 1) / kSlotThreads;
  if (grid > need) grid = need;
  OGJK_CK(cudaMemsetAsync(ticket, 0, sizeof(unsigned), t_stream));
  kern<<<(unsigned)grid, kSlotThreads, smem, t_stream>>>(c1, c2, nv1, nv2, simp, dist, (unsigned)n, utab, ticket, 0u, pairs);
  return finish_launch(packed ? "gjk slots kernel (SoA-4 packed pool)" : "gjk slots kernel");
}

// tickets per atomic of the ws loader on dense batches (>= 32); development override OGJK_WS_CHUNK
unsigned ws_dense_chunk() {
  const char* e = getenv("OGJK_WS_CHUNK");
  const int v = e ? atoi(e) : 64;
  return v < 32 ? 32u : (unsigned)v;
}

// warp-specialised slot kernel.  Configurations (compute warps, lanes per pair): (8,1) when 256 slots fit an SM;
// otherwise (8,2) -- 128 slots, two lanes per pair -- or (4,1); (2,1) = 64 slots for 65..~140 vertices per body.  normals/queue/count non-null = fused EPA gate.
// development override: OGJK_WS_LP=1|2 picks between (4,1) and (8,2) for the 128-slot case.
int ws_config(int nv1, int nv2, int* lp, int esize) {
  const size_t sb = slot_bytes(nv1, nv2, esize);
  *lp = 1;
  if (ws_fixed_bytes(256, esize) + 256 * sb + kSlotPadBytes <= 227u * 1024u) return 8;
  const char* e = getenv("OGJK_WS_LP");
  if (e && atoi(e) == 2 && esize == 4 &&
      ws_fixed_bytes(128) + 128 * (size_t)ws_slot_layout(nv1, nv2, 2).stride + kSlotPadBytes <= 227u * 1024u) {
    *lp = 2;
    return 8;
  }
  if (ws_fixed_bytes(128, esize) + 128 * sb + kSlotPadBytes <= 227u * 1024u) return 4;
  if (ws_fixed_bytes(64, esize) + 64 * sb + kSlotPadBytes <= 227u * 1024u) return 2;  // fp32: 65..~140 vertices per body
  // (one compute warp, 32 slots, up to ~270 vertices per body, was measured too: it loses to gjk_uniform_kernel,
  //  2.0e8 against 3.1e8 pairs/s at 256 vertices, while two compute warps win, 7.3e8 against 4.7e8 at 96 vertices)
  return 0;
}
int ws_compute_warps(int nv1, int nv2, int esize) {
  int lp;
  return ws_config(nv1, nv2, &lp, esize);
}
template <typename T, int CW, int LP>
int launch_gjk_slots_ws_cw(int n, int nv1, const T* c1, int nv2, const T* c2, SimplexT<T>* simp, T* dist, T* nrm,
                           int* queue, int* count, const CollisionPair* pairs) {
  const uint16_t* utab = nullptr;
  if (int rc = device_unified_table(&utab)) return rc;
  unsigned* ticket = nullptr;
  if (int rc = ticket_buffer(&ticket)) return rc;
  constexpr int nslots = CW * 32 / LP;
  constexpr int es = (int)sizeof(T);
  const size_t smem = (size_t)ws_fixed_bytes(nslots, es) + kSlotPadBytes + (size_t)nslots * ws_slot_layout(nv1, nv2, LP, es).stride;
  constexpr int threads = (CW + 2) * 32;
  constexpr bool eq_ok = LP == 1 && es == 4;  // the interleaved two-body scan exists for fp32 only
  auto kern = pairs ? gjk_slots_ws_kernel<T, CW, LP, eq_ok, true>  // one pool: equal vertex counts by construction
                    : (eq_ok && nv1 == nv2) ? gjk_slots_ws_kernel<T, CW, LP, eq_ok, false>
                                            : gjk_slots_ws_kernel<T, CW, LP, false, false>;
  long long grid = 0;
  if (int rc = persistent_grid(kern, threads, smem, &grid)) return rc;
  const long long need = ((long long)n + nslots - 1) / nslots;
  if (grid > need) grid = need;
  OGJK_CK(cudaMemsetAsync(ticket, 0, sizeof(unsigned), t_stream));
  kern<<<(unsigned)grid, threads, smem, t_stream>>>(c1, c2, nv1, nv2, simp, dist, (unsigned)n, utab, ticket, 0u,
                                                    nrm, queue, count, pairs, ws_dense_chunk());
  return finish_launch("gjk slots (warp-specialised) kernel");
}
// ---- fp16 pre-scan slot kernel (gjk_slots16.cuh): fp32 batches whose bodies have 8 * NB vertices, NB in 4..8 ---------
// Bit-exact and tested, but NOT a default: measured on B200 (profiles/r2_experiments.txt, section D) it runs config 2 in
// 1.07 ms against 0.77 ms for the fp32-slot kernel -- twice the slots, but the conversion and the exact verification
// of the candidates cost more issue slots than the cheaper scan saves.  OGJK_GJK_KERNEL=slots16 selects it for every
// supported shape (=slotsws32 names the fp32-slot kernel explicitly); OGJK_S16_CFG=<K> (development) picks another
// depth of the fetch ring for the 64+64-vertex dense instance.
bool slots16_shape(int nv1, int nv2) { return nv1 == nv2 && nv1 % 8 == 0 && nv1 >= 32 && nv1 <= 64; }
bool use_slots16(int nv1, int nv2, int esize) {
  return esize == 4 && slots16_shape(nv1, nv2) && forced_kernel() == 5;
}
template <int NB, bool IDX, int K>
int launch_gjk_slots16_inst(int n, const float* c1, const float* c2, SimplexT<float>* simp, float* dist, float* nrm,
                            int* queue, int* count, const CollisionPair* pairs) {
  const uint16_t* utab = nullptr;
  if (int rc = device_unified_table(&utab)) return rc;
  unsigned* ticket = nullptr;
  if (int rc = ticket_buffer(&ticket)) return rc;
  constexpr size_t smem = s16_smem_bytes(NB, NB);
  static_assert(smem <= 227u * 1024u, "slots do not fit");
  auto kern = gjk_slots16_kernel<NB, NB, IDX, K>;
  long long grid = 0;
  if (int rc = persistent_grid(kern, kS16Threads, smem, &grid)) return rc;
  const long long need = ((long long)n + kS16Slots - 1) / kS16Slots;
  if (grid > need) grid = need;
  OGJK_CK(cudaMemsetAsync(ticket, 0, sizeof(unsigned), t_stream));
  kern<<<(unsigned)grid, kS16Threads, smem, t_stream>>>(c1, c2, simp, dist, (unsigned)n, utab, ticket, nrm, queue, count, pairs);
  return finish_launch("gjk slots (fp16 pre-scan) kernel");
}
template <int NB>
int launch_gjk_slots16_nb(int n, const float* c1, const float* c2, SimplexT<float>* simp, float* dist, float* nrm,
                          int* queue, int* count, const CollisionPair* pairs) {
  if (pairs) return launch_gjk_slots16_inst<NB, true, 4>(n, c1, c2, simp, dist, nrm, queue, count, pairs);
  if constexpr (NB == 8) {
    const char* e = getenv("OGJK_S16_CFG");  // development: pairs a warp keeps in flight
    const int cfg = e ? atoi(e) : 0;
    if (cfg == 2) return launch_gjk_slots16_inst<NB, false, 2>(n, c1, c2, simp, dist, nrm, queue, count, pairs);
    if (cfg == 3) return launch_gjk_slots16_inst<NB, false, 3>(n, c1, c2, simp, dist, nrm, queue, count, pairs);
    if (cfg == 5) return launch_gjk_slots16_inst<NB, false, 5>(n, c1, c2, simp, dist, nrm, queue, count, pairs);
  }
  return launch_gjk_slots16_inst<NB, false, 4>(n, c1, c2, simp, dist, nrm, queue, count, pairs);
}
int launch_gjk_slots16(int n, int nv, const float* c1, const float* c2, SimplexT<float>* simp, float* dist, float* nrm,
                       int* queue, int* count, const CollisionPair* pairs) {
  switch (nv / 8) {
    case 4: return launch_gjk_slots16_nb<4>(n, c1, c2, simp, dist, nrm, queue, count, pairs);
    case 5: return launch_gjk_slots16_nb<5>(n, c1, c2, simp, dist, nrm, queue, count, pairs);
    case 6: return launch_gjk_slots16_nb<6>(n, c1, c2, simp, dist, nrm, queue, count, pairs);
    case 7: return launch_gjk_slots16_nb<7>(n, c1, c2, simp, dist, nrm, queue, count, pairs);
    default: return launch_gjk_slots16_nb<8>(n, c1, c2, simp, dist, nrm, queue, count, pairs);
  }
}

template <typename T>
int launch_gjk_slots_ws(int n, int nv1, const T* c1, int nv2, const T* c2, SimplexT<T>* simp, T* dist, T* nrm,
                        int* queue, int* count, const CollisionPair* pairs = nullptr) {
  if constexpr (sizeof(T) == 4) {
    if (use_slots16(nv1, nv2, 4)) return launch_gjk_slots16(n, nv1, c1, c2, simp, dist, nrm, queue, count, pairs);
  }
  int lp = 1;
  const int cw = ws_config(nv1, nv2, &lp, (int)sizeof(T));
  if (cw == 8 && lp == 1) return launch_gjk_slots_ws_cw<T, 8, 1>(n, nv1, c1, nv2, c2, simp, dist, nrm, queue, count, pairs);
  if constexpr (sizeof(T) == 4) {
    if (cw == 8 && lp == 2)
      return launch_gjk_slots_ws_cw<T, 8, 2>(n, nv1, c1, nv2, c2, simp, dist, nrm, queue, count, pairs);
  }
  if (cw == 4) return launch_gjk_slots_ws_cw<T, 4, 1>(n, nv1, c1, nv2, c2, simp, dist, nrm, queue, count, pairs);
  if (cw == 2) return launch_gjk_slots_ws_cw<T, 2, 1>(n, nv1, c1, nv2, c2, simp, dist, nrm, queue, count, pairs);
  return 1;
}

// policy: the warp-specialised kernel pays off when slots are so large that the self-service kernel would be down
// to one 4-warp CTA per SM (measured on B200: 64+64 vertices 1.28e9 vs 1.15e9 pairs/s; at 32+32 vertices, where
// the self-service kernel runs 8 warps per SM, it is the other way round: 1.6e9 vs 2.6e9)
bool use_ws_kernel(int nv1, int nv2, int esize) {
  const char* e = getenv("OGJK_WS_MIN_SLOT");  // development override (bytes)
  const int thr = e ? atoi(e) : 880;
  if (ws_compute_warps(nv1, nv2, esize) == 0) return false;
  // fp64 (measured on B200, 1 Mi pairs): the self-service kernel wins wherever its 128 slots fit -- 32+32 vertices
  // 1.06 ms against 1.45 ms, 16+16 0.53 against 0.78 -- so the warp-specialised kernel (64 slots) only takes over
  // above that: 64+64 vertices 1.99 ms against 2.14 ms for the general kernel
  if (esize == 8 && !e)
    return (size_t)kSlotFixedBytes + kSlotPadBytes + (size_t)kSlotThreads * slot_bytes(nv1, nv2, esize) > 227u * 1024u;
  return (int)slot_bytes(nv1, nv2, esize) >= thr;
}

template <typename T>
int launch_gjk_slots_if(int n, int nv1, const T* c1, int nv2, const T* c2, SimplexT<T>* simp, T* dist) {
  constexpr int es = (int)sizeof(T);
  const int force = forced_kernel();
  if (force == 2 || force == 3) return 1;
  if (force >= 4 || (force == 0 && n >= 32768 && use_ws_kernel(nv1, nv2, es)))
    return launch_gjk_slots_ws<T>(n, nv1, c1, nv2, c2, simp, dist, nullptr, nullptr, nullptr);
  if ((size_t)kSlotFixedBytes + kSlotPadBytes + (size_t)kSlotThreads * slot_bytes(nv1, nv2, es) > 227u * 1024u) return 1;
  if (force == 0 && n < 32768) return 1;
  return launch_gjk_slots<T>(n, nv1, c1, nv2, c2, simp, dist);
}

template <typename T>
int launch_gjk_uniform(int n, int nv1, const T* c1, int nv2, const T* c2, SimplexT<T>* simp, T* dist) {
  const int nv = nv1 > nv2 ? nv1 : nv2;
  const bool aligned = (((uintptr_t)c1 | (uintptr_t)c2) & 15u) == 0 && nv1 % 4 == 0 && nv2 % 4 == 0;
  if (!aligned || nv > 256) return 1;
  if (forced_kernel() == 3) return 1;
  const uint32_t* tabs = nullptr;
  if (int rc = device_tables(&tabs)) return rc < 0 ? rc : -1;
  {
    const int rc = launch_gjk_slots_if<T>(n, nv1, c1, nv2, c2, simp, dist);
    if (rc <= 0) return rc;
  }
  // measured on B200 (profiles/): the register-resident kernel wins for fp32 with 17..256 vertices; for tiny
  // polytopes and for fp64 (twice the registers per vertex) the general kernel is faster.
  if constexpr (sizeof(T) == 4) {
    if (nv <= 16) return 1;
    else if (nv <= 32) launch_uniform_instance<T, 4, 8>(c1, c2, nv1, nv2, simp, dist, n, tabs);
    else if (nv <= 64) launch_uniform_instance<T, 8, 8>(c1, c2, nv1, nv2, simp, dist, n, tabs);
    else if (nv <= 128) launch_uniform_instance<T, 16, 8>(c1, c2, nv1, nv2, simp, dist, n, tabs);
    else launch_uniform_instance<T, 32, 8>(c1, c2, nv1, nv2, simp, dist, n, tabs);
  } else {
    return 1;
  }
  return finish_launch("gjk uniform kernel");
}

// ---- scratch for the EPA work queue ------------------------------------------------------------------------------
// counters: [0] queued pairs, [1] ticket of the first EPA kernel, [2] overflow count, [3] ticket of the overflow pass
struct EpaQueue {
  int* counters = nullptr;
  int* queue = nullptr;     // n entries: colliding pairs, appended by the gate
  int* overflow = nullptr;  // n entries: pairs the small-work-area kernel handed back
};
int epa_queue_buffers(size_t n, EpaQueue* q) {
  StreamScratch* ss = nullptr;
  if (int rc = stream_scratch(&ss)) return rc;
  int* p = nullptr;
  if (int rc = scratch_grow(ss->epa, 2 * n + 4, &p)) return rc;
  q->counters = p;
  q->queue = p + 4;
  q->overflow = p + 4 + n;
  OGJK_CK(cudaMemsetAsync(p, 0, 4 * sizeof(int), t_stream));
  return 0;
}

template <typename T, typename Source>
int launch_epa_warp_queue(const Source& src, int n, SimplexT<T>* d_simplices, T* d_distances, T* d_normals,
                          const int* queue, int* counters, int sms) {
  constexpr int wpb = EpaConfig<T>::kWarpsPerBlock;
  int per_sm = 0;
  OGJK_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, epa_queue_kernel<T, Source>, wpb * 32, 0));
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)sms * per_sm;
  const long long need = ((long long)n + wpb - 1) / wpb;
  if (grid > need) grid = need;
  epa_queue_kernel<T, Source><<<(unsigned)grid, wpb * 32, 0, t_stream>>>(src, d_simplices, d_distances, d_normals, queue,
                                                                        counters);
  return finish_launch("epa kernel");
}

// BYREGS: the register budget MINB is registers per thread (epa_group_kernel_regs) instead of resident warps per SM
template <typename T, int G, int KV, typename WT, int MINB, int WPC, typename Source, bool BYREGS = false>
int launch_epa_group(const Source& src, int n, SimplexT<T>* d_simplices, T* d_distances, T* d_normals, const EpaQueue& q,
                     int sms) {
  constexpr int threads = 32 * WPC;
  constexpr int groups = threads / G;
  const size_t smem = (size_t)groups * sizeof(WT);
  void (*kern)(const Source, SimplexT<T>*, T*, T*, const int*, int*, int*, int, int) = nullptr;
  if constexpr (BYREGS) kern = epa_group_kernel_regs<T, G, KV, WT, MINB, WPC, Source>;
  else kern = epa_group_kernel<T, G, KV, WT, MINB, WPC, Source>;
  long long grid = 0;
  if (int rc = persistent_grid(kern, threads, smem, &grid)) return rc;
  const long long need = ((long long)n + groups - 1) / groups;
  if (grid > need) grid = need;
  const bool sync_saved = t_sync;
  if (WT::kSmall) t_sync = false;  // the overflow pass follows
  // service batching of the group kernel (epa_group.cuh): development override OGJK_EPA_SVC=<batch><defer>, e.g. 23
  int svc_batch = 2, svc_defer = 4;
  if (const char* e = getenv("OGJK_EPA_SVC")) {
    const int v = atoi(e);
    svc_batch = v / 10 > 0 ? v / 10 : 1;
    svc_defer = v % 10;
  }
  kern<<<(unsigned)grid, threads, smem, t_stream>>>(src, d_simplices, d_distances, d_normals, q.queue, q.counters, q.overflow,
                                                    svc_batch, svc_defer);
  int rc = finish_launch("epa group kernel");
  t_sync = sync_saved;
  if (rc || !WT::kSmall) return rc;
  // pairs that did not fit the small work area (a fraction of a percent): warp-per-pair kernel, full-size area
  return launch_epa_warp_queue<T, Source>(src, n, d_simplices, d_distances, d_normals, q.overflow, q.counters + 2, sms);
}

// Persistent EPA over a device-side queue of colliding pairs.  Default (measured on B200, profiles/r2_experiments.txt
// sections A, E, G): bodies of up to 32 vertices take the sub-warp group kernel, G = 4 lanes per pair, followed by the
// overflow pass -- config 3: 4.88 ms against 7.64 ms per Mi pairs for one warp per pair, config 5: 8.40 against 14.1 ms
// per 3.96 M pairs, 16-vertex bodies 3.43 against 6.43; larger bodies (the support scan grows, the bookkeeping does
// not) take one warp per pair -- config 2 (64 vertices, shallow contacts): 0.30 ms against 0.34 (G = 8) and 0.48
// (G = 4); only deep 64-vertex contacts favour G = 8 (3.04 against 3.42 ms per 512 Ki pairs).  Measured and dropped
// (profiles/r2y5_ab_epa_svc.txt): G = 8 with four cached vertices per lane and the 1.4 KB area at 20 / 24 warps per SM
// -- config 3 5.50 / 5.48 ms.  Development overrides: OGJK_EPA_KERNEL=warp|group (full-size area, 8 lanes)|small4|small8,
// OGJK_EPA_SVC=<batch><defer>, OGJK_EPA_AREA=small (the 1.7 KB area at 128 registers instead of the 1.4 KB ones at 96).
template <typename T, typename Source>
int launch_epa_queue(const Source& src, int n, int nv_hint, SimplexT<T>* d_simplices, T* d_distances, T* d_normals,
                     const EpaQueue& q) {
  int dev = 0, sms = 0;
  OGJK_CK(cudaGetDevice(&dev));
  OGJK_CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const char* e = getenv("OGJK_EPA_KERNEL");
  int mode = nv_hint <= 32 ? 4 : 0;  // 0 warp per pair, 4 / 8 small work area with that many lanes, 1 full-size group
  if (e) mode = !strcmp(e, "warp") ? 0 : !strcmp(e, "group") ? 1 : !strcmp(e, "small4") ? 4 : !strcmp(e, "small8") ? 8 : mode;
  constexpr int minb = sizeof(T) == 4 ? 16 : 8;
  if (mode == 4) {
    // Occupancy decides here (the kernel is latency-bound: issue slots 63 % busy at 16 warps per SM), and the register file
    // only knows two useful budgets -- each scheduler owns 16 K registers, so 128 registers per thread = 4 warps per
    // scheduler and 96 = 5; anything in between still runs 16 warps per SM (measured: profiles/r2y6_ab_epa_svc.txt).
    //   fp32, bodies of up to 16 vertices (KV = 4 fits 96 registers): 1.4 KB area, five 4-warp CTAs = 20 warps per SM,
    //     3.86 -> 3.43 ms per Mi pairs;
    //   fp32, up to 32 vertices: the LEAN instantiation (per-pair state in the work area, epa_group.cuh) at 96 registers,
    //     20 warps per SM: config 3 5.09 -> 4.88 ms, config 5 8.89 -> 8.40 ms per 3.96 M pairs (profiles/r2y7_ab_epa_svc.txt);
    //   otherwise (fp64, OGJK_EPA_AREA=small): 1.7 KB area at 128 registers, eight 2-warp CTAs = 16 warps per SM (the 1 KB
    //     of shared memory that the system reserves per CTA is paid 8 times instead of 15: config 3 5.31 -> 5.07 ms).
    const char* area = getenv("OGJK_EPA_AREA");  // development: small
    const bool small_area = area && !strcmp(area, "small");
    if constexpr (sizeof(T) == 4) {
      if (!small_area) {
        if (nv_hint <= 16)
          return launch_epa_group<T, 4, 4, EpaWorkTiny<T>, 20, 4, Source>(src, n, d_simplices, d_distances, d_normals, q, sms);
        return launch_epa_group<T, 4, 8, EpaWorkLean<T>, 96, 4, Source, true>(src, n, d_simplices, d_distances, d_normals, q, sms);
      }
    }
    if (nv_hint <= 16)
      return launch_epa_group<T, 4, 4, EpaWorkSmall<T>, minb, 2, Source>(src, n, d_simplices, d_distances, d_normals, q, sms);
    return launch_epa_group<T, 4, 8, EpaWorkSmall<T>, minb, 2, Source>(src, n, d_simplices, d_distances, d_normals, q, sms);
  }
  if (mode == 8) return launch_epa_group<T, 8, 8, EpaWorkSmall<T>, minb, 1, Source>(src, n, d_simplices, d_distances, d_normals, q, sms);
  if (mode == 1) return launch_epa_group<T, 8, 8, EpaWork<T>, 1, 1, Source>(src, n, d_simplices, d_distances, d_normals, q, sms);
  return launch_epa_warp_queue<T, Source>(src, n, d_simplices, d_distances, d_normals, q.queue, q.counters, sms);
}

// EPA launch: small batches get one warp per pair; large ones go through gate + compaction + a persistent
// queue kernel so that only colliding pairs occupy warps.
template <typename T, typename Source>
int launch_epa(const Source& src, int n, int nv_hint, SimplexT<T>* d_simplices, T* d_distances, T* d_normals) {
  if (!d_normals) return fail_msg("contact_normals must not be NULL on the device path");
  constexpr int wpb = EpaConfig<T>::kWarpsPerBlock;
  if (n < 8192) {
    const unsigned grid = (unsigned)(((long long)n + wpb - 1) / wpb);
    epa_kernel<T, Source><<<grid, wpb * 32, 0, t_stream>>>(src, d_simplices, d_distances, d_normals, n);
    return finish_launch("epa kernel");
  }
  EpaQueue q;
  if (int rc = epa_queue_buffers((size_t)n, &q)) return rc;
  epa_gate_kernel<T><<<(unsigned)(((long long)n + 255) / 256), 256, 0, t_stream>>>(d_simplices, d_distances, d_normals, n,
                                                                            q.queue, q.counters);
  ++t_launches;
  OGJK_CK(cudaGetLastError());
  return launch_epa_queue<T, Source>(src, n, nv_hint, d_simplices, d_distances, d_normals, q);
}

// GJK then EPA on a dense uniform device batch.  When the warp-specialised slot kernel applies, its finisher warp
// also does the EPA gate (normals of separated pairs, queue of colliding ones), which saves the gate kernel's pass
// over distances + witnesses; otherwise the two stages run back to back as in the reference (openGJK.cu:2854-2883).
int stage_mark(int which) {
  if (!t_timing) return 0;
  if (which == 0) {
    if (t_stage_used == t_stage_pool.size()) {
      StageEvents se;
      for (auto& ev : se.e) OGJK_CK(cudaEventCreate(&ev));
      t_stage_pool.push_back(se);
    }
    ++t_stage_used;
  }
  OGJK_CK(cudaEventRecord(t_stage_pool[t_stage_used - 1].e[which], t_stream));
  return 0;
}

template <typename T>
int launch_gjk_epa_uniform(int n, int nv1, const T* c1, int nv2, const T* c2, SimplexT<T>* simp, T* dist, T* nrm) {
  if (!nrm) return fail_msg("contact_normals must not be NULL on the device path");
  UniformSource<T> src{c1, c2, nv1, nv2};
  if (int rc = stage_mark(0)) return rc;
  {
    const bool aligned = (((uintptr_t)c1 | (uintptr_t)c2) & 15u) == 0 && nv1 % 4 == 0 && nv2 % 4 == 0;
    const int force = forced_kernel();
    if (aligned && n >= 32768 && (force == 0 || force >= 4) && use_ws_kernel(nv1, nv2, (int)sizeof(T))) {
      EpaQueue q;
      if (int rc = epa_queue_buffers((size_t)n, &q)) return rc;
      const bool sync_saved = t_sync;
      t_sync = false;  // no need to synchronise between the two stages
      int rc = launch_gjk_slots_ws<T>(n, nv1, c1, nv2, c2, simp, dist, nrm, q.queue, q.counters);
      t_sync = sync_saved;
      if (!rc) rc = stage_mark(1);
      if (!rc) rc = launch_epa_queue<T, UniformSource<T>>(src, n, nv1 > nv2 ? nv1 : nv2, simp, dist, nrm, q);
      if (!rc) rc = stage_mark(2);
      return rc;
    }
  }
  const bool sync_saved = t_sync;
  t_sync = false;
  int rc = launch_gjk_uniform<T>(n, nv1, c1, nv2, c2, simp, dist);
  if (rc > 0) rc = launch_gjk_generic<T>(src, n, (nv1 + nv2) / 2, simp, dist);
  t_sync = sync_saved;
  if (!rc) rc = stage_mark(1);
  if (!rc) rc = launch_epa<T>(src, n, nv1 > nv2 ? nv1 : nv2, simp, dist, nrm);
  if (!rc) rc = stage_mark(2);
  return rc;
}

// ---- indexed batches over a uniform pool ------------------------------------------------------------------------
// Pools uploaded by this library (allocate_indexed_device, the host-level *_indexed calls) are remembered when every
// polytope has the same vertex count (a multiple of 4): their coordinates then lie densely in the blob, pair t's two
// bodies are pool[idx1] and pool[idx2], and the slot kernels can be fed from the gkCollisionPair list.  Descriptor
// arrays built by the caller (the visualiser does that) are not in the registry and take the general kernel.
constexpr int kGjkStage = 1, kEpaStage = 2;  // = Stage::kGjk, Stage::kEpa below
struct PoolInfo {
  const void* coords;
  int nv;      // > 0: uniform pool, dense coordinates; 0: only `max_nv` is known
  int count;
  int max_nv;  // vertex-count hint for the general kernel's lane selection
};
std::mutex g_pool_mutex;
std::vector<std::pair<const void*, PoolInfo>> g_pools;  // keyed by the device descriptor pointer
void register_pool(const void* d_desc, const void* d_coords, int nv, int count, int max_nv = 0) {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  const PoolInfo info{d_coords, nv, count, max_nv > 0 ? max_nv : nv};
  for (auto& p : g_pools)
    if (p.first == d_desc) {  // a recycled address (the caller cudaFree'd a registered array): the new entry wins
      p.second = info;
      return;
    }
  g_pools.push_back({d_desc, info});
}
void unregister_pool(const void* d_desc) {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  for (size_t i = 0; i < g_pools.size(); ++i)
    if (g_pools[i].first == d_desc) {
      g_pools.erase(g_pools.begin() + i);
      return;
    }
}
bool lookup_pool(const void* d_desc, PoolInfo* out) {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  for (const auto& p : g_pools)
    if (p.first == d_desc) {
      *out = p.second;
      return true;
    }
  return false;
}

// The registry is only a hint: before a fast kernel reads the remembered layout, the live descriptors are checked
// against it on the device (one pass over `count` descriptors, a 4-byte read-back).  *ok = false sends the call to
// the descriptor-reading general kernels, so a caller that edits d_bd[i].coord / numpoints after upload, or re-uses
// a freed address without telling the library, still gets results for the descriptors it passed.
// One or two descriptor arrays per call share one flag, one read-back (into pinned host memory) and one synchronisation:
// on small batches the check is most of what the call costs.
thread_local int* t_pinned_flag = nullptr;
template <typename T>
int pool_layouts_hold(const void* d_desc1, const PoolInfo& info1, const void* d_desc2, const PoolInfo* info2, int count,
                      bool* ok) {
  *ok = false;
  if (info1.nv <= 0 || count > info1.count) return 0;
  if (d_desc2 && (info2->nv <= 0 || count > info2->count)) return 0;
  StreamScratch* ss = nullptr;
  if (int rc = stream_scratch(&ss)) return rc;
  int* flag = nullptr;
  if (int rc = scratch_grow(ss->flag, 1, &flag)) return rc;
  if (!t_pinned_flag) OGJK_CK(cudaMallocHost(&t_pinned_flag, sizeof(int)));
  OGJK_CK(cudaMemsetAsync(flag, 0, sizeof(int), t_stream));
  const unsigned grid = (unsigned)((count + 255) / 256);
  validate_dense_kernel<T><<<grid, 256, 0, t_stream>>>((const PolytopeT<T>*)d_desc1, count, (const T*)info1.coords, info1.nv, flag);
  ++t_launches;
  if (d_desc2) {
    validate_dense_kernel<T><<<grid, 256, 0, t_stream>>>((const PolytopeT<T>*)d_desc2, count, (const T*)info2->coords, info2->nv, flag);
    ++t_launches;
  }
  OGJK_CK(cudaGetLastError());
  *t_pinned_flag = 1;
  OGJK_CK(cudaMemcpyAsync(t_pinned_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, t_stream));
  OGJK_CK(cudaStreamSynchronize(t_stream));
  *ok = *t_pinned_flag == 0;
  return 0;
}
template <typename T>
int pool_layout_holds(const void* d_desc, const PoolInfo& info, int count, bool* ok) {
  return pool_layouts_hold<T>(d_desc, info, nullptr, nullptr, count, ok);
}

// GJK (and optionally the fused EPA gate + EPA) over `pairs` into a uniform fp32 pool.  Returns 1 when the batch does
// not qualify for the slot kernels.
template <typename T>
int launch_indexed_uniform(int n, const PoolInfo& pool, const CollisionPair* d_pairs, const PolytopeT<T>* d_desc,
                           SimplexT<T>* simp, T* dist, T* nrm, int stages) {
  constexpr int es = (int)sizeof(T);
  const int nv = pool.nv;
  const T* base = (const T*)pool.coords;
  const int force = forced_kernel();
  if (nv <= 0 || nv % 4 || n < 32768 || force == 2 || force == 3 || !(stages & kGjkStage)) return 1;
  const bool ws = force >= 4 || (force == 0 && use_ws_kernel(nv, nv, es));
  const bool v2_fits = (size_t)kSlotFixedBytes + kSlotPadBytes + (size_t)kSlotThreads * slot_bytes(nv, nv, es) <= 227u * 1024u;
  if (!ws && !v2_fits) return 1;
  if (ws && ws_compute_warps(nv, nv, es) == 0) return 1;
  const bool epa = (stages & kEpaStage) != 0;
  if (!epa) {
    return ws ? launch_gjk_slots_ws<T>(n, nv, base, nv, base, simp, dist, nullptr, nullptr, nullptr, d_pairs)
              : launch_gjk_slots<T>(n, nv, base, nv, base, simp, dist, d_pairs, pool.count);
  }
  if (!nrm) return fail_msg("contact_normals must not be NULL on the device path");
  IndexedSource<T> src{d_desc, d_pairs};
  const bool sync_saved = t_sync;
  t_sync = false;
  int rc = stage_mark(0);
  if (rc) {
    t_sync = sync_saved;
    return rc;
  }
  if (ws) {  // gate fused into the finisher warp
    EpaQueue q;
    if ((rc = epa_queue_buffers((size_t)n, &q))) {
      t_sync = sync_saved;
      return rc;
    }
    rc = launch_gjk_slots_ws<T>(n, nv, base, nv, base, simp, dist, nrm, q.queue, q.counters, d_pairs);
    t_sync = sync_saved;
    if (!rc) rc = stage_mark(1);
    if (!rc) rc = launch_epa_queue<T, IndexedSource<T>>(src, n, nv, simp, dist, nrm, q);
    if (!rc) rc = stage_mark(2);
    return rc;
  }
  rc = launch_gjk_slots<T>(n, nv, base, nv, base, simp, dist, d_pairs, pool.count);
  t_sync = sync_saved;
  if (!rc) rc = stage_mark(1);
  if (!rc) rc = launch_epa<T>(src, n, nv, simp, dist, nrm);
  if (!rc) rc = stage_mark(2);
  return rc;
}

// numpoints of the first descriptor of a device array (vertex-count hint for lane selection)
template <typename T>
int peek_numpoints(const PolytopeT<T>* d_bd, int* nv) {
  int v = 0;
  OGJK_CK(cudaMemcpyAsync(&v, &d_bd->numpoints, sizeof(int), cudaMemcpyDeviceToHost, t_stream));
  OGJK_CK(cudaStreamSynchronize(t_stream));
  *nv = v;
  return 0;
}

// ---- host-side flattening ----------------------------------------------------------------------------
// One device blob per descriptor array: coordinates of polytope i start at a 16-byte aligned offset so that
// vector loads and bulk copies stay legal; the descriptor's `coord` is re-pointed into the blob, as the
// reference does (openGJK.cu:2928-2943).
template <typename T>
struct Flattened {
  PolytopeT<T>* d_desc = nullptr;
  T* d_coord = nullptr;
  long long max_nv = 0;
  int uniform_nv = 0;  // > 0: every polytope has this many vertices (a multiple of 4), coordinates dense in the blob
};

template <typename T>
bool dense_uniform_range(const PolytopeT<T>* bd, size_t lo, size_t hi);  // below

template <typename T>
int flatten_upload(int n, const PolytopeT<T>* bd, Flattened<T>& out) {
  size_t total = 0;
  long long max_nv = 0;
  bool uniform = n > 0 && bd[0].numpoints % 4 == 0;
  const size_t align = 16 / sizeof(T);
  for (int i = 0; i < n; ++i) {
    if (bd[i].numpoints < 1 || !bd[i].coord) return fail_msg("polytope with no vertices");
    total += ((size_t)bd[i].numpoints * 3 + align - 1) / align * align;
    if (bd[i].numpoints > max_nv) max_nv = bd[i].numpoints;
    if (bd[i].numpoints != bd[0].numpoints) uniform = false;
  }
  out.uniform_nv = uniform ? bd[0].numpoints : 0;
  out.max_nv = max_nv;
  if (uniform && dense_uniform_range(bd, 0, (size_t)n)) {
    // Uniform batch whose coordinates already lie back to back in the caller's memory (what every driver of the
    // reference builds): no staging copy -- one DMA straight from the caller's array -- and the descriptors are
    // written on the device instead of being staged and uploaded (the reference stages both, openGJK.cu:2917-2948).
    cudaError_t e;
    if ((e = cudaMalloc(&out.d_coord, total * sizeof(T))) != cudaSuccess) return fail("cudaMalloc(coords)", e);
    if ((e = cudaMalloc(&out.d_desc, (size_t)n * sizeof(PolytopeT<T>))) != cudaSuccess) {
      cudaFree(out.d_coord);
      out.d_coord = nullptr;
      return fail("cudaMalloc(desc)", e);
    }
    init_polytopes_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, t_stream>>>(out.d_desc, out.d_coord, nullptr, nullptr,
                                                                            bd[0].numpoints, n);
    ++t_launches;
    e = cudaMemcpyAsync(out.d_coord, bd[0].coord, total * sizeof(T), cudaMemcpyHostToDevice, t_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(t_stream);
    if (e != cudaSuccess) {
      cudaFree(out.d_coord);
      cudaFree(out.d_desc);
      out.d_coord = nullptr;
      out.d_desc = nullptr;
      return fail("H2D coords", e);
    }
    return 0;
  }
  T* h_coord = nullptr;
  PolytopeT<T>* h_desc = nullptr;
  OGJK_CK(cudaMallocHost(&h_coord, total * sizeof(T)));
  cudaError_t e = cudaMallocHost(&h_desc, (size_t)n * sizeof(PolytopeT<T>));
  if (e != cudaSuccess) {
    cudaFreeHost(h_coord);
    return fail("cudaMallocHost", e);
  }
  int rc = 0;
  do {
    if ((e = cudaMalloc(&out.d_coord, total * sizeof(T))) != cudaSuccess) { rc = fail("cudaMalloc(coords)", e); break; }
    if ((e = cudaMalloc(&out.d_desc, (size_t)n * sizeof(PolytopeT<T>))) != cudaSuccess) { rc = fail("cudaMalloc(desc)", e); break; }
    size_t off = 0;
    for (int i = 0; i < n; ++i) {
      const size_t cnt = (size_t)bd[i].numpoints * 3;
      memcpy(h_coord + off, bd[i].coord, cnt * sizeof(T));
      h_desc[i] = bd[i];
      h_desc[i].coord = out.d_coord + off;
      const size_t padded = (cnt + align - 1) / align * align;
      for (size_t k = cnt; k < padded; ++k) h_coord[off + k] = T(0);
      off += padded;
    }
    if ((e = cudaMemcpyAsync(out.d_coord, h_coord, total * sizeof(T), cudaMemcpyHostToDevice, t_stream)) != cudaSuccess) { rc = fail("H2D coords", e); break; }
    if ((e = cudaMemcpyAsync(out.d_desc, h_desc, (size_t)n * sizeof(PolytopeT<T>), cudaMemcpyHostToDevice, t_stream)) != cudaSuccess) { rc = fail("H2D desc", e); break; }
    if ((e = cudaStreamSynchronize(t_stream)) != cudaSuccess) { rc = fail("H2D sync", e); break; }
  } while (0);
  cudaFreeHost(h_coord);
  cudaFreeHost(h_desc);
  if (rc) {
    cudaFree(out.d_coord);
    cudaFree(out.d_desc);
    out.d_coord = nullptr;
    out.d_desc = nullptr;
  }
  out.max_nv = max_nv;
  return rc;
}

template <typename T>
void release(Flattened<T>& f) {
  cudaFree(f.d_coord);
  cudaFree(f.d_desc);
  f.d_coord = nullptr;
  f.d_desc = nullptr;
}

struct SyncOverride {  // high-level calls copy results back right after the launches: no need to sync in between
  bool saved;
  SyncOverride() : saved(t_sync) { t_sync = false; }
  ~SyncOverride() { t_sync = saved; }
};

// ---- high-level implementations ----------------------------------------------------------------------
enum Stage : int { kGjk = 1, kEpa = 2 };

// ---- cached device buffers + helper streams for the host-pointer fast path -----------------------------------
// The reference allocates and frees six device buffers on every high-level call (openGJK.cu:2889-2954, 3034-3048).
// Here they are grow-only per (thread, device) and re-used, so a steady-state call costs no cudaMalloc/cudaFree.
enum Slot : int { kSlotC1 = 0, kSlotC2, kSlotSimp, kSlotDist, kSlotNrm, kSlotPairs, kSlotCount };
struct DevicePool {
  void* ptr[kSlotCount] = {};
  size_t cap[kSlotCount] = {};
  cudaStream_t s_copy = nullptr, s_comp = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_done;
};
thread_local DevicePool t_pool[kMaxDevices];

int pool_get(DevicePool& p, int slot, size_t bytes, void** out) {
  if (p.cap[slot] < bytes) {
    if (p.ptr[slot]) cudaFree(p.ptr[slot]);
    p.ptr[slot] = nullptr;
    p.cap[slot] = 0;
    OGJK_CK(cudaMalloc(&p.ptr[slot], bytes));
    p.cap[slot] = bytes;
  }
  *out = p.ptr[slot];
  return 0;
}
int pool_streams(DevicePool& p, size_t chunks) {
  if (!p.s_copy) {
    OGJK_CK(cudaStreamCreateWithFlags(&p.s_copy, cudaStreamNonBlocking));
    OGJK_CK(cudaStreamCreateWithFlags(&p.s_comp, cudaStreamNonBlocking));
    OGJK_CK(cudaStreamCreateWithFlags(&p.s_out, cudaStreamNonBlocking));
  }
  while (p.ev_in.size() < chunks) {
    cudaEvent_t a, b;
    OGJK_CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    OGJK_CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    p.ev_in.push_back(a);
    p.ev_done.push_back(b);
  }
  return 0;
}

// Is the descriptor array a dense uniform batch, i.e. numpoints identical and coord[i] = coord[0] + i*nv*3 ?
// Checked for descriptors [lo, hi) against descriptor 0: the host path validates chunk by chunk, right before it
// queues the chunk's copies, so that walking 64 MB of descriptors hides behind the PCIe transfer instead of preceding it.
template <typename T>
bool dense_uniform_range(const PolytopeT<T>* bd, size_t lo, size_t hi) {
  const int nv = bd[0].numpoints;
  const T* base = bd[0].coord;
  if (nv < 1 || !base) return false;
  const size_t stride = (size_t)nv * 3;
  for (size_t i = lo; i < hi; ++i)
    if (bd[i].numpoints != nv || bd[i].coord != base + i * stride) return false;
  return true;
}

struct StreamOverride {  // run the launch helpers on one of the pool's streams
  cudaStream_t saved;
  explicit StreamOverride(cudaStream_t s) : saved(t_stream) { t_stream = s; }
  ~StreamOverride() { t_stream = saved; }
};

// Host-pointer fast path for dense uniform batches: the caller's coordinate arrays are DMA'd as they are (no host
// staging copy, no descriptor upload), in chunks, with the kernels of chunk k overlapping the upload of chunk k+1
// and the download of chunk k-1.  Returns 1 if the batch does not qualify.
template <typename T>
int run_pairs_host_dense(int n, const PolytopeT<T>* bd1, const PolytopeT<T>* bd2, SimplexT<T>* simplices,
                         T* distances, T* normals, int stages) {
  // a cheap look at both ends decides whether to try; every chunk is validated in full before it is queued
  const size_t probe = n < 64 ? (size_t)n : 64;
  if (!dense_uniform_range(bd1, 0, probe) || !dense_uniform_range(bd2, 0, probe) ||
      !dense_uniform_range(bd1, (size_t)n - probe, (size_t)n) || !dense_uniform_range(bd2, (size_t)n - probe, (size_t)n))
    return 1;
  const int nv1 = bd1[0].numpoints, nv2 = bd2[0].numpoints;
  if (nv1 % 4 || nv2 % 4 || nv1 > 256 || nv2 > 256) return 1;
  if ((((uintptr_t)bd1[0].coord | (uintptr_t)bd2[0].coord) & 15u) != 0) return 1;
  int dev = 0;
  OGJK_CK(cudaGetDevice(&dev));
  DevicePool& P = t_pool[dev];
  T *d_c1, *d_c2, *d_dist, *d_nrm = nullptr;
  SimplexT<T>* d_simp;
  if (int rc = pool_get(P, kSlotC1, (size_t)n * nv1 * 3 * sizeof(T), (void**)&d_c1)) return rc;
  if (int rc = pool_get(P, kSlotC2, (size_t)n * nv2 * 3 * sizeof(T), (void**)&d_c2)) return rc;
  if (int rc = pool_get(P, kSlotSimp, (size_t)n * sizeof(SimplexT<T>), (void**)&d_simp)) return rc;
  if (int rc = pool_get(P, kSlotDist, (size_t)n * sizeof(T), (void**)&d_dist)) return rc;
  if (stages & kEpa)
    if (int rc = pool_get(P, kSlotNrm, (size_t)n * 3 * sizeof(T), (void**)&d_nrm)) return rc;

  // Chunks are sized in PAIRS: at least 65536, so that every chunk is large enough for the persistent slot kernels
  // (>= 32768 pairs, ~3.5 waves of 148 CTAs x 128 slots) -- the host-pointer API then runs the same kernels as the
  // device-resident path -- and at least ~24 MB of coordinates, so that small polytopes still move in large DMAs.
  const size_t pair_bytes = (size_t)(nv1 + nv2) * 3 * sizeof(T);
  size_t chunk_pairs = (size_t)(24u << 20) / pair_bytes;
  if (chunk_pairs < 65536) chunk_pairs = 65536;
  const size_t chunks = ((size_t)n + chunk_pairs - 1) / chunk_pairs;
  chunk_pairs = (((size_t)n + chunks - 1) / chunks + 255) & ~(size_t)255;  // equal chunks: no undersized tail
  if (int rc = pool_streams(P, chunks)) return rc;
  // EPA-only calls overwrite their inputs (simplices, distances): there the whole descriptor array is validated before
  // anything is queued, so that a batch that turns out not to be dense is handed to the general path untouched.  For
  // calls that start with GJK the outputs are pure outputs, and validation runs chunk by chunk behind the transfers.
  if (!(stages & kGjk) && (!dense_uniform_range(bd1, 0, (size_t)n) || !dense_uniform_range(bd2, 0, (size_t)n))) return 1;
  // order the pool's streams after whatever the caller queued on the selected stream
  OGJK_CK(cudaEventRecord(P.ev_done[0], t_stream));
  OGJK_CK(cudaStreamWaitEvent(P.s_copy, P.ev_done[0], 0));
  OGJK_CK(cudaStreamWaitEvent(P.s_comp, P.ev_done[0], 0));
  OGJK_CK(cudaStreamWaitEvent(P.s_out, P.ev_done[0], 0));
  // every exit below -- error or not -- leaves no copy to or from the caller's arrays in flight
  struct Drain {
    DevicePool& p;
    ~Drain() {
      cudaStreamSynchronize(p.s_copy);
      cudaStreamSynchronize(p.s_comp);
      cudaStreamSynchronize(p.s_out);
    }
  } drain{P};

  const T* h_c1 = bd1[0].coord;
  const T* h_c2 = bd2[0].coord;
  SyncOverride nosync;
  for (size_t k = 0; k < chunks; ++k) {
    const size_t lo = k * chunk_pairs;
    if (lo >= (size_t)n) break;
    const int m = (int)(((size_t)n - lo) < chunk_pairs ? ((size_t)n - lo) : chunk_pairs);
    if ((stages & kGjk) && (!dense_uniform_range(bd1, lo, lo + m) || !dense_uniform_range(bd2, lo, lo + m)))
      return 1;  // not a dense batch after all: the general path redoes the whole call (queued work is drained)
    OGJK_CK(cudaMemcpyAsync(d_c1 + lo * nv1 * 3, h_c1 + lo * nv1 * 3, (size_t)m * nv1 * 3 * sizeof(T),
                            cudaMemcpyHostToDevice, P.s_copy));
    OGJK_CK(cudaMemcpyAsync(d_c2 + lo * nv2 * 3, h_c2 + lo * nv2 * 3, (size_t)m * nv2 * 3 * sizeof(T),
                            cudaMemcpyHostToDevice, P.s_copy));
    if (!(stages & kGjk)) {  // EPA only: the caller's GJK results are inputs (examples/gpu/example.cu:104-105)
      OGJK_CK(cudaMemcpyAsync(d_simp + lo, simplices + lo, (size_t)m * sizeof(SimplexT<T>), cudaMemcpyHostToDevice,
                              P.s_copy));
      OGJK_CK(cudaMemcpyAsync(d_dist + lo, distances + lo, (size_t)m * sizeof(T), cudaMemcpyHostToDevice, P.s_copy));
    }
    OGJK_CK(cudaEventRecord(P.ev_in[k], P.s_copy));
    OGJK_CK(cudaStreamWaitEvent(P.s_comp, P.ev_in[k], 0));
    {
      StreamOverride on(P.s_comp);
      if ((stages & kGjk) && (stages & kEpa)) {
        OGJK_CK(cudaMemsetAsync(d_nrm + lo * 3, 0, (size_t)m * 3 * sizeof(T), P.s_comp));
        if (int rc = launch_gjk_epa_uniform<T>(m, nv1, d_c1 + lo * nv1 * 3, nv2, d_c2 + lo * nv2 * 3, d_simp + lo,
                                               d_dist + lo, d_nrm + lo * 3))
          return rc;
      } else if (stages & kGjk) {
        int rc = launch_gjk_uniform<T>(m, nv1, d_c1 + lo * nv1 * 3, nv2, d_c2 + lo * nv2 * 3, d_simp + lo, d_dist + lo);
        if (rc > 0) {
          UniformSource<T> src{d_c1 + lo * nv1 * 3, d_c2 + lo * nv2 * 3, nv1, nv2};
          rc = launch_gjk_generic<T>(src, m, (nv1 + nv2) / 2, d_simp + lo, d_dist + lo);
        }
        if (rc) return rc;
      }
      if ((stages & kEpa) && !(stages & kGjk)) {
        OGJK_CK(cudaMemsetAsync(d_nrm + lo * 3, 0, (size_t)m * 3 * sizeof(T), P.s_comp));
        UniformSource<T> src{d_c1 + lo * nv1 * 3, d_c2 + lo * nv2 * 3, nv1, nv2};
        if (int rc = launch_epa<T>(src, m, nv1 > nv2 ? nv1 : nv2, d_simp + lo, d_dist + lo, d_nrm + lo * 3)) return rc;
      }
    }
    OGJK_CK(cudaEventRecord(P.ev_done[k], P.s_comp));
    OGJK_CK(cudaStreamWaitEvent(P.s_out, P.ev_done[k], 0));
    OGJK_CK(cudaMemcpyAsync(simplices + lo, d_simp + lo, (size_t)m * sizeof(SimplexT<T>), cudaMemcpyDeviceToHost, P.s_out));
    OGJK_CK(cudaMemcpyAsync(distances + lo, d_dist + lo, (size_t)m * sizeof(T), cudaMemcpyDeviceToHost, P.s_out));
    if ((stages & kEpa) && normals)
      OGJK_CK(cudaMemcpyAsync(normals + lo * 3, d_nrm + lo * 3, (size_t)m * 3 * sizeof(T), cudaMemcpyDeviceToHost, P.s_out));
  }
  OGJK_CK(cudaStreamSynchronize(P.s_out));
  OGJK_CK(cudaStreamSynchronize(P.s_comp));
  return 0;
}

template <typename T>
int run_pairs_host(int n, const PolytopeT<T>* bd1, const PolytopeT<T>* bd2, SimplexT<T>* simplices, T* distances,
                   T* normals, T* witness1, T* witness2, int stages) {
  if (n <= 0) return 0;
  if (!bd1 || !bd2 || !simplices || !distances) return fail_msg("null argument");
  {
    const int rc = run_pairs_host_dense<T>(n, bd1, bd2, simplices, distances, normals, stages);
    if (rc <= 0) {
      if (rc == 0 && (witness1 || witness2))
        for (int i = 0; i < n; ++i)
          for (int c = 0; c < 3; ++c) {
            if (witness1) witness1[3 * (size_t)i + c] = simplices[i].witnesses[0][c];
            if (witness2) witness2[3 * (size_t)i + c] = simplices[i].witnesses[1][c];
          }
      return rc;
    }
  }
  Flattened<T> f1, f2;
  SimplexT<T>* d_simp = nullptr;
  T* d_dist = nullptr;
  T* d_nrm = nullptr;
  int rc = 0;
  cudaError_t e;
  do {
    if ((rc = flatten_upload(n, bd1, f1))) break;
    if ((rc = flatten_upload(n, bd2, f2))) break;
    if ((e = cudaMalloc(&d_simp, (size_t)n * sizeof(SimplexT<T>))) != cudaSuccess) { rc = fail("cudaMalloc(simplices)", e); break; }
    if ((e = cudaMalloc(&d_dist, (size_t)n * sizeof(T))) != cudaSuccess) { rc = fail("cudaMalloc(distances)", e); break; }
    if (stages & kEpa) {
      if ((e = cudaMalloc(&d_nrm, (size_t)n * 3 * sizeof(T))) != cudaSuccess) { rc = fail("cudaMalloc(normals)", e); break; }
      if ((e = cudaMemsetAsync(d_nrm, 0, (size_t)n * 3 * sizeof(T), t_stream)) != cudaSuccess) { rc = fail("memset", e); break; }
    }
    if (stages & kGjk) {
      if ((e = cudaMemsetAsync(d_simp, 0, (size_t)n * sizeof(SimplexT<T>), t_stream)) != cudaSuccess) { rc = fail("memset", e); break; }
    } else {  // EPA only: the caller's GJK results are the input (reference examples/gpu/example.cu:104-105)
      if ((e = cudaMemcpyAsync(d_simp, simplices, (size_t)n * sizeof(SimplexT<T>), cudaMemcpyHostToDevice, t_stream)) != cudaSuccess) { rc = fail("H2D simplices", e); break; }
      if ((e = cudaMemcpyAsync(d_dist, distances, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, t_stream)) != cudaSuccess) { rc = fail("H2D distances", e); break; }
    }
    {
      SyncOverride nosync;
      DescSource<T> src{f1.d_desc, f2.d_desc};
      const int hint = (int)((f1.max_nv + f2.max_nv) / 2);
      if ((stages & kGjk) && (rc = launch_gjk_generic<T>(src, n, hint, d_simp, d_dist))) break;
      if ((stages & kEpa) && (rc = launch_epa<T>(src, n, (int)(f1.max_nv > f2.max_nv ? f1.max_nv : f2.max_nv), d_simp, d_dist, d_nrm))) break;
    }
    if ((e = cudaMemcpyAsync(distances, d_dist, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, t_stream)) != cudaSuccess) { rc = fail("D2H distances", e); break; }
    if ((e = cudaMemcpyAsync(simplices, d_simp, (size_t)n * sizeof(SimplexT<T>), cudaMemcpyDeviceToHost, t_stream)) != cudaSuccess) { rc = fail("D2H simplices", e); break; }
    if ((stages & kEpa) && normals) {
      if ((e = cudaMemcpyAsync(normals, d_nrm, (size_t)n * 3 * sizeof(T), cudaMemcpyDeviceToHost, t_stream)) != cudaSuccess) { rc = fail("D2H normals", e); break; }
    }
    if ((e = cudaStreamSynchronize(t_stream)) != cudaSuccess) { rc = fail("sync", e); break; }
    if (witness1 || witness2) {
      for (int i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) {
          if (witness1) witness1[3 * (size_t)i + c] = simplices[i].witnesses[0][c];
          if (witness2) witness2[3 * (size_t)i + c] = simplices[i].witnesses[1][c];
        }
    }
  } while (0);
  release(f1);
  release(f2);
  cudaFree(d_simp);
  cudaFree(d_dist);
  cudaFree(d_nrm);
  return rc;
}

template <typename T>
int run_indexed_host(int num_polytopes, int num_pairs, const PolytopeT<T>* polytopes, const CollisionPair* pairs,
                     SimplexT<T>* simplices, T* distances, T* normals, int stages) {
  // Host-level indexed call (reference openGJK.cu:3274-3311: upload pool + pair list, launch, copy everything back).
  // Here the pair list is processed in chunks of ~1 Mi pairs through pooled device buffers: the pair records of chunk
  // k+1 go up and the results of chunk k-1 come down (124 bytes per pair -- the dominant transfer of this call) while
  // the kernels of chunk k run; nothing is allocated per call except the (small) pool.
  if (num_pairs <= 0 || num_polytopes <= 0) return 0;
  if (!polytopes || !pairs || !simplices || !distances) return fail_msg("null argument");
  const size_t np = (size_t)num_pairs;
  int dev = 0;
  OGJK_CK(cudaGetDevice(&dev));
  DevicePool& P = t_pool[dev];
  CollisionPair* d_pairs = nullptr;
  SimplexT<T>* d_simp = nullptr;
  T* d_dist = nullptr;
  T* d_nrm = nullptr;
  if (int rc = pool_get(P, kSlotPairs, np * sizeof(CollisionPair), (void**)&d_pairs)) return rc;
  if (int rc = pool_get(P, kSlotSimp, np * sizeof(SimplexT<T>), (void**)&d_simp)) return rc;
  if (int rc = pool_get(P, kSlotDist, np * sizeof(T), (void**)&d_dist)) return rc;
  if (stages & kEpa)
    if (int rc = pool_get(P, kSlotNrm, np * 3 * sizeof(T), (void**)&d_nrm)) return rc;
  size_t chunks = (np + (1u << 20) - 1) >> 20;
  size_t chunk_pairs = ((np + chunks - 1) / chunks + 255) & ~(size_t)255;  // equal chunks, all large enough for the slot kernels
  if (int rc = pool_streams(P, chunks)) return rc;
  Flattened<T> pool;
  if (int rc = flatten_upload(num_polytopes, polytopes, pool)) {
    release(pool);
    return rc;
  }
  int rc = 0;
  {
    // order the pool's streams after the upload of the polytope pool (queued on the selected stream)
    cudaEventRecord(P.ev_done[0], t_stream);
    cudaStreamWaitEvent(P.s_copy, P.ev_done[0], 0);
    cudaStreamWaitEvent(P.s_comp, P.ev_done[0], 0);
    cudaStreamWaitEvent(P.s_out, P.ev_done[0], 0);
    struct Drain {  // every exit leaves no copy to or from the caller's arrays in flight
      DevicePool& p;
      ~Drain() {
        cudaStreamSynchronize(p.s_copy);
        cudaStreamSynchronize(p.s_comp);
        cudaStreamSynchronize(p.s_out);
      }
    } drain{P};
    SyncOverride nosync;
    const PoolInfo info{pool.d_coord, pool.uniform_nv, num_polytopes, (int)pool.max_nv};
    for (size_t k = 0; k < chunks && !rc; ++k) {
      const size_t lo = k * chunk_pairs;
      if (lo >= np) break;
      const size_t m = np - lo < chunk_pairs ? np - lo : chunk_pairs;
      for (size_t i = lo; i < lo + m; ++i)
        if (pairs[i].idx1 < 0 || pairs[i].idx1 >= num_polytopes || pairs[i].idx2 < 0 || pairs[i].idx2 >= num_polytopes) {
          rc = fail_msg("pair index out of range");
          break;
        }
      if (rc) break;
      cudaError_t e = cudaMemcpyAsync(d_pairs + lo, pairs + lo, m * sizeof(CollisionPair), cudaMemcpyHostToDevice, P.s_copy);
      if (e == cudaSuccess && !(stages & kGjk)) {  // EPA alone: simplices and distances are inputs
        e = cudaMemcpyAsync(d_simp + lo, simplices + lo, m * sizeof(SimplexT<T>), cudaMemcpyHostToDevice, P.s_copy);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_dist + lo, distances + lo, m * sizeof(T), cudaMemcpyHostToDevice, P.s_copy);
      }
      if (e == cudaSuccess) e = cudaEventRecord(P.ev_in[k], P.s_copy);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(P.s_comp, P.ev_in[k], 0);
      if (e != cudaSuccess) {
        rc = fail("H2D pairs", e);
        break;
      }
      {
        StreamOverride on(P.s_comp);
        if (stages & kEpa) cudaMemsetAsync(d_nrm + 3 * lo, 0, m * 3 * sizeof(T), P.s_comp);
        if (stages & kGjk) cudaMemsetAsync(d_simp + lo, 0, m * sizeof(SimplexT<T>), P.s_comp);
        IndexedSource<T> src{pool.d_desc, d_pairs + lo};
        rc = launch_indexed_uniform<T>((int)m, info, d_pairs + lo, pool.d_desc, d_simp + lo, d_dist + lo,
                                       d_nrm ? d_nrm + 3 * lo : nullptr, stages);
        if (rc == 1) {  // not a uniform fp32 pool / small batch: the general kernels
          rc = 0;
          if (stages & kGjk) rc = launch_gjk_generic<T>(src, (int)m, (int)pool.max_nv, d_simp + lo, d_dist + lo);
          if (!rc && (stages & kEpa)) rc = launch_epa<T>(src, (int)m, (int)pool.max_nv, d_simp + lo, d_dist + lo, d_nrm + 3 * lo);
        }
      }
      if (rc) break;
      e = cudaEventRecord(P.ev_done[k], P.s_comp);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(P.s_out, P.ev_done[k], 0);
      if (e == cudaSuccess) e = cudaMemcpyAsync(simplices + lo, d_simp + lo, m * sizeof(SimplexT<T>), cudaMemcpyDeviceToHost, P.s_out);
      if (e == cudaSuccess) e = cudaMemcpyAsync(distances + lo, d_dist + lo, m * sizeof(T), cudaMemcpyDeviceToHost, P.s_out);
      if (e == cudaSuccess && (stages & kEpa) && normals)
        e = cudaMemcpyAsync(normals + 3 * lo, d_nrm + 3 * lo, m * 3 * sizeof(T), cudaMemcpyDeviceToHost, P.s_out);
      if (e != cudaSuccess) rc = fail("D2H results", e);
    }
    if (!rc) {
      cudaError_t e = cudaStreamSynchronize(P.s_out);
      if (e == cudaSuccess) e = cudaStreamSynchronize(P.s_comp);
      if (e != cudaSuccess) rc = fail("indexed host call", e);
    }
  }
  release(pool);
  return rc;
}

// ---- multi-GPU fan-out of the host-pointer calls -------------------------------------------
// The reference's API has no device argument and uses the current device.  When more than one device is selected
// (ogjk_set_devices(), or OGJK_DEVICES=all|<count>|<i,j,...> in the environment) the host-pointer entry points split
// the pair range into one contiguous slice per device: a persistent host thread per device runs the single-device
// path on its slice -- own streams, own cached device buffers, chunked H2D / kernels / D2H overlapped as above -- and
// writes straight into the caller's arrays at the slice offset.  Pairs are independent, so nothing is exchanged;
// for the indexed calls the polytope pool is uploaded to every device and only the pair list is sliced.
class DeviceWorker {
 public:
  explicit DeviceWorker(int dev) : dev_(dev), th_([this] { loop(); }) {}
  void submit(std::function<int()> job) {
    std::lock_guard<std::mutex> lk(m_);
    job_ = std::move(job);
    state_ = 1;
    cv_.notify_all();
  }
  int wait(std::string* err) {
    std::unique_lock<std::mutex> lk(m_);
    cv_.wait(lk, [&] { return state_ == 2; });
    state_ = 0;
    *err = err_;
    return rc_;
  }

 private:
  void loop() {
    const cudaError_t e = cudaSetDevice(dev_);
    for (;;) {
      std::unique_lock<std::mutex> lk(m_);
      cv_.wait(lk, [&] { return state_ == 1; });
      std::function<int()> job = std::move(job_);
      lk.unlock();
      int rc = e == cudaSuccess ? job() : fail("cudaSetDevice", e);
      const std::string msg = rc ? t_err : std::string();
      lk.lock();
      rc_ = rc;
      err_ = msg;
      state_ = 2;
      cv_.notify_all();
    }
  }
  int dev_;
  std::mutex m_;
  std::condition_variable cv_;
  std::function<int()> job_;
  int state_ = 0;  // 0 idle, 1 job posted, 2 result ready
  int rc_ = 0;
  std::string err_;
  std::thread th_;  // last member: starts after the others are constructed; never joined (workers live to exit)
};

std::mutex g_dev_mutex;
std::vector<int> g_devices;       // selected devices; empty or one entry = single-device behaviour
bool g_devices_from_env = false;  // environment consulted
DeviceWorker* g_workers[kMaxDevices] = {};

std::vector<int> selected_devices() {
  std::lock_guard<std::mutex> lk(g_dev_mutex);
  if (!g_devices_from_env) {
    g_devices_from_env = true;
    const char* e = getenv("OGJK_DEVICES");
    if (e && *e && g_devices.empty()) {
      int count = 0;
      cudaGetDeviceCount(&count);
      if (!strcmp(e, "all")) {
        for (int i = 0; i < count; ++i) g_devices.push_back(i);
      } else if (strchr(e, ',')) {
        for (const char* q = e; *q;) {
          const int d = atoi(q);
          if (d >= 0 && d < count) g_devices.push_back(d);
          q = strchr(q, ',');
          if (!q) break;
          ++q;
        }
      } else {
        const int want = atoi(e);
        for (int i = 0; i < want && i < count; ++i) g_devices.push_back(i);
      }
    }
  }
  return g_devices;
}

// runs part(g, G) for g in [0, G) on the workers of `devs`; first failure wins
int fan_out(const std::vector<int>& devs, const std::function<int(int, int)>& part) {
  const int G = (int)devs.size();
  {
    std::lock_guard<std::mutex> lk(g_dev_mutex);
    for (int d : devs)
      if (!g_workers[d]) g_workers[d] = new DeviceWorker(d);
  }
  // one fan-out at a time per process: the workers (and their cached buffers) are shared
  static std::mutex fan_mutex;
  std::lock_guard<std::mutex> fan(fan_mutex);
  int rc = 0;
  std::string first;
  int pending[kMaxDevices];  // slice a worker is busy with, or -1 (a device listed twice runs its slices back to back)
  for (int& x : pending) x = -1;
  auto collect = [&](int dev) {
    std::string msg;
    const int r = g_workers[dev]->wait(&msg);
    if (r && !rc) {
      rc = r;
      first = "device " + std::to_string(dev) + ": " + msg;
    }
    pending[dev] = -1;
  };
  for (int g = 0; g < G; ++g) {
    if (pending[devs[g]] >= 0) collect(devs[g]);
    g_workers[devs[g]]->submit([=] { return part(g, G); });
    pending[devs[g]] = g;
  }
  for (int d = 0; d < kMaxDevices; ++d)
    if (pending[d] >= 0) collect(d);
  if (rc) t_err = first;
  return rc;
}

constexpr long long kMinPairsPerDevice = 65536;  // below this a second device costs more than it saves

template <typename T>
int run_pairs_host_multi(int n, const PolytopeT<T>* bd1, const PolytopeT<T>* bd2, SimplexT<T>* simplices, T* distances,
                         T* normals, T* witness1, T* witness2, int stages) {
  std::vector<int> devs = selected_devices();
  if (n > 0 && devs.size() > 1) {
    const long long fit = (long long)n / kMinPairsPerDevice;
    if (fit < (long long)devs.size()) devs.resize(fit < 1 ? 1 : (size_t)fit);
  }
  if (devs.size() <= 1 || n <= 0 || !bd1 || !bd2 || !simplices || !distances)
    return run_pairs_host<T>(n, bd1, bd2, simplices, distances, normals, witness1, witness2, stages);
  return fan_out(devs, [=](int g, int G) {
    const long long lo = (long long)n * g / G, hi = (long long)n * (g + 1) / G;
    return run_pairs_host<T>((int)(hi - lo), bd1 + lo, bd2 + lo, simplices + lo, distances + lo,
                             normals ? normals + 3 * lo : nullptr, witness1 ? witness1 + 3 * lo : nullptr,
                             witness2 ? witness2 + 3 * lo : nullptr, stages);
  });
}

template <typename T>
int run_indexed_host_multi(int num_polytopes, int num_pairs, const PolytopeT<T>* polytopes, const CollisionPair* pairs,
                           SimplexT<T>* simplices, T* distances, T* normals, int stages) {
  std::vector<int> devs = selected_devices();
  if (num_pairs > 0 && devs.size() > 1) {
    const long long fit = (long long)num_pairs / kMinPairsPerDevice;
    if (fit < (long long)devs.size()) devs.resize(fit < 1 ? 1 : (size_t)fit);
  }
  if (devs.size() <= 1 || num_pairs <= 0 || num_polytopes <= 0 || !polytopes || !pairs || !simplices || !distances)
    return run_indexed_host<T>(num_polytopes, num_pairs, polytopes, pairs, simplices, distances, normals, stages);
  return fan_out(devs, [=](int g, int G) {
    const long long lo = (long long)num_pairs * g / G, hi = (long long)num_pairs * (g + 1) / G;
    return run_indexed_host<T>(num_polytopes, (int)(hi - lo), polytopes, pairs + lo, simplices + lo, distances + lo,
                               normals ? normals + 3 * lo : nullptr, stages);
  });
}

// ---- broad phase (reference visualization/integrate_final_gjk.cu:916-1002) -------------
int broadphase_pairs(int n, const float4* d_pos, float cell_size, float boundary, int grid_size, CollisionPair* d_pairs,
                     int max_pairs, long long* num_pairs) {
  if (num_pairs) *num_pairs = 0;
  if (n <= 0) return 0;
  if (!d_pos || (!d_pairs && max_pairs > 0)) return fail_msg("null argument");
  if (!(cell_size > 0.0f) || grid_size < 1 || grid_size > 512) return fail_msg("bad grid (cell_size > 0, 1 <= grid_size <= 512)");
  const size_t cells = (size_t)grid_size * grid_size * grid_size;
  // obj_cell[n] | cell_count[cells] | cell_start[cells + 1] | cell_objs[n] | pair_counts[n] | pair_offsets[n + 1] |
  // obj_id[n] | sorted keys[n] | radix-sort temporary
  int key_bits = 1;
  while ((1ll << key_bits) < (long long)cells) ++key_bits;
  size_t sort_bytes = 0;
  OGJK_CK(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const int*)nullptr, (int*)nullptr, (const int*)nullptr,
                                          (int*)nullptr, n, 0, key_bits, t_stream));
  const size_t ints = 6 * (size_t)n + 2 * cells + 2 + (sort_bytes + 3) / 4 + 4;
  StreamScratch* ss = nullptr;
  if (int rc = stream_scratch(&ss)) return rc;
  int* obj_cell = nullptr;
  if (int rc = scratch_grow(ss->bp, ints, &obj_cell)) return rc;
  int* cell_count = obj_cell + n;
  int* cell_start = cell_count + cells;
  int* cell_objs = cell_start + cells + 1;
  int* pair_counts = cell_objs + n;
  int* pair_offsets = pair_counts + n;
  int* obj_id = pair_offsets + n + 1;
  int* keys_sorted = obj_id + n;
  void* sort_tmp = (void*)(((uintptr_t)(keys_sorted + n) + 15u) & ~(uintptr_t)15u);
  OGJK_CK(cudaMemsetAsync(cell_count, 0, cells * sizeof(int), t_stream));
  const unsigned tb = (unsigned)((n + 255) / 256);
  const unsigned wb = (unsigned)(((long long)n * 32 + 255) / 256);
  bp_histogram_kernel<<<tb, 256, 0, t_stream>>>(d_pos, n, cell_size, boundary, grid_size, obj_cell, obj_id, cell_count);
  bp_exclusive_scan_kernel<<<1, 1024, 0, t_stream>>>(cell_count, cell_start, (int)cells);
  // cell lists in ascending object id: stable sort of the ids by cell (deterministic, unlike an atomic cursor fill)
  OGJK_CK(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, obj_cell, keys_sorted, obj_id, cell_objs, n, 0, key_bits,
                                          t_stream));
  bp_pairs_kernel<false><<<wb, 256, 0, t_stream>>>(d_pos, n, cell_size, boundary, grid_size, cell_start, cell_objs,
                                                   pair_counts, nullptr, nullptr, 0);
  bp_exclusive_scan_kernel<<<1, 1024, 0, t_stream>>>(pair_counts, pair_offsets, n);
  t_launches += 5;
  OGJK_CK(cudaGetLastError());
  int total = 0;
  OGJK_CK(cudaMemcpyAsync(&total, pair_offsets + n, sizeof(int), cudaMemcpyDeviceToHost, t_stream));
  OGJK_CK(cudaStreamSynchronize(t_stream));
  if (num_pairs) *num_pairs = total;
  if (total > 0 && max_pairs > 0) {
    bp_pairs_kernel<true><<<wb, 256, 0, t_stream>>>(d_pos, n, cell_size, boundary, grid_size, cell_start, cell_objs,
                                                    nullptr, pair_offsets, d_pairs, max_pairs);
    return finish_launch("broad phase");
  }
  return 0;
}

// ---- contact response (reference visualization/integrate_final_gjk.cu:572-689, 1039-1054) ---
template <typename T>
int contact_response(int num_pairs, const CollisionPair* d_pairs, const T* d_dist, const SimplexT<T>* d_simp,
                     const T* d_nrm, const int* d_sub_mesh_body, int num_objects, float4* d_pos, const float4* d_vel_ping,
                     float4* d_vel_pong, const float4* d_ang_ping, float4* d_ang_pong, const float4* d_quats,
                     const float* d_inv_inertia, ContactParams prm) {
  if (num_objects <= 0) return 0;
  if (num_pairs < 0) num_pairs = 0;
  if (!d_pos || !d_vel_ping || !d_vel_pong || !d_ang_ping || !d_ang_pong || !d_quats || !d_inv_inertia)
    return fail_msg("null argument");
  if (num_pairs > 0 && (!d_pairs || !d_dist || !d_simp || !d_nrm)) return fail_msg("null argument");
  if (d_vel_ping == d_vel_pong || d_ang_ping == d_ang_pong) return fail_msg("ping and pong buffers must differ");
  if (num_pairs > (1 << 30)) return fail_msg("too many pairs");
  const size_t slots = 2 * (size_t)num_pairs;
  int key_bits = 1;
  while ((1ll << key_bits) <= (long long)num_objects) ++key_bits;  // the sentinel key is num_objects itself
  size_t sort_bytes = 0;
  if (slots)
    OGJK_CK(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned*)nullptr, (unsigned*)nullptr,
                                            (const unsigned*)nullptr, (unsigned*)nullptr, (int)slots, 0, key_bits,
                                            t_stream));
  // counts[num_objects + 1] | seg_start[num_objects + 2] | keys[slots] x 2 | slot ids[slots] x 2 | positions_out | sort temp
  const size_t head = 2 * (size_t)num_objects + 4;
  const size_t ints = head + 4 * slots + 4 * (size_t)num_objects + (sort_bytes + 3) / 4 + 8;
  StreamScratch* ss = nullptr;
  if (int rc = stream_scratch(&ss)) return rc;
  int* counts = nullptr;
  if (int rc = scratch_grow(ss->cr, ints, &counts)) return rc;
  int* seg_start = counts + num_objects + 1;
  unsigned* keys_in = (unsigned*)(counts + ((head + 3) & ~(size_t)3));
  unsigned* keys_out = keys_in + slots;
  unsigned* slots_in = keys_out + slots;
  unsigned* slots_out = slots_in + slots;
  float4* pos_out = (float4*)(((uintptr_t)(slots_out + slots) + 15u) & ~(uintptr_t)15u);
  void* sort_tmp = (void*)(pos_out + num_objects);
  OGJK_CK(cudaMemsetAsync(counts, 0, ((size_t)num_objects + 1) * sizeof(int), t_stream));
  if (num_pairs > 0) {
    cr_keys_kernel<T><<<(unsigned)((num_pairs + 255) / 256), 256, 0, t_stream>>>(
        d_pairs, d_dist, d_sub_mesh_body, prm.epsilon, num_pairs, num_objects, keys_in, slots_in, counts);
    ++t_launches;
    OGJK_CK(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, keys_in, keys_out, slots_in, slots_out, (int)slots, 0,
                                            key_bits, t_stream));
  }
  bp_exclusive_scan_kernel<<<1, 1024, 0, t_stream>>>(counts, seg_start, num_objects);
  ++t_launches;
  cr_accumulate_kernel<T><<<(unsigned)(((long long)num_objects * 32 + 255) / 256), 256, 0, t_stream>>>(
      d_pos, pos_out, d_vel_ping, d_vel_pong, d_ang_ping, d_ang_pong, d_quats, d_inv_inertia, d_pairs, d_dist, d_simp,
      d_nrm, d_sub_mesh_body, prm, num_objects, seg_start, slots_out);
  OGJK_CK(cudaMemcpyAsync(d_pos, pos_out, (size_t)num_objects * sizeof(float4), cudaMemcpyDeviceToDevice, t_stream));
  return finish_launch("contact response");
}

}  // namespace

// =======================================================================================================
extern "C" {

const char* ogjk_last_error(void) { return t_err.c_str(); }
const char* ogjk_version(void) { return "opengjk-b200 0.1 (sm_100a)"; }
int ogjk_selected_device_count(void) { return (int)selected_devices().size(); }
int ogjk_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return -1;
  return n;
}
int ogjk_set_device(int device) {
  OGJK_CK(cudaSetDevice(device));
  return 0;
}
int ogjk_set_stream(void* stream) {
  t_stream = (cudaStream_t)stream;
  return 0;
}
int ogjk_set_sync(int enabled) {
  t_sync = enabled != 0;
  return 0;
}
int ogjk_set_timing(int enabled) {
  t_timing = enabled != 0;
  t_stage_used = 0;
  return 0;
}
int ogjk_stage_times(double* gjk_ms, double* epa_ms, int* calls) {
  double g = 0, e = 0;
  for (size_t i = 0; i < t_stage_used; ++i) {
    float a = 0, b = 0;
    OGJK_CK(cudaEventSynchronize(t_stage_pool[i].e[2]));
    OGJK_CK(cudaEventElapsedTime(&a, t_stage_pool[i].e[0], t_stage_pool[i].e[1]));
    OGJK_CK(cudaEventElapsedTime(&b, t_stage_pool[i].e[1], t_stage_pool[i].e[2]));
    g += a;
    e += b;
  }
  if (gjk_ms) *gjk_ms = g;
  if (epa_ms) *epa_ms = e;
  if (calls) *calls = (int)t_stage_used;
  t_stage_used = 0;
  return 0;
}
int ogjk_transform_to_world_device(int num_submeshes, const float* d_positions, const float* d_quats,
                                   const float* d_scales, const float* d_verts_local, float* d_verts_world,
                                   const int* d_vert_offsets, const int* d_vert_counts, const int* d_sub_mesh_body,
                                   int uniform_count) {
  if (num_submeshes <= 0) return 0;
  if (!d_positions || !d_quats || !d_scales || !d_verts_local || !d_verts_world) return fail_msg("null argument");
  if ((!d_vert_offsets || !d_vert_counts) && uniform_count < 1) return fail_msg("vertex layout missing");
  transform_to_world_kernel<<<(unsigned)num_submeshes, 64, 0, t_stream>>>(
      (const float4*)d_positions, (const float4*)d_quats, d_scales, d_verts_local, d_verts_world, d_vert_offsets,
      d_vert_counts, d_sub_mesh_body, uniform_count, num_submeshes);
  return finish_launch("transform_to_world");
}
int ogjk_release_pool(const void* d_polytopes) {
  unregister_pool(d_polytopes);
  return 0;
}
int ogjk_broadphase_pairs_device(int num_objects, const float* d_pos_radius, float cell_size, float boundary,
                                 int grid_size, void* d_pairs, int max_pairs, long long* num_pairs) {
  return broadphase_pairs(num_objects, (const float4*)d_pos_radius, cell_size, boundary, grid_size,
                          (CollisionPair*)d_pairs, max_pairs, num_pairs);
}
int ogjk_set_devices(int count, const int* devices) {
  int have = 0;
  OGJK_CK(cudaGetDeviceCount(&have));
  std::vector<int> v;
  for (int i = 0; i < count; ++i) {
    const int d = devices ? devices[i] : i;
    if (d < 0 || d >= have || d >= kMaxDevices) return fail_msg("ogjk_set_devices: device ordinal out of range");
    v.push_back(d);  // a device may be listed more than once: its slices then run back to back
  }
  std::lock_guard<std::mutex> lk(g_dev_mutex);
  g_devices = v;
  g_devices_from_env = true;  // an explicit selection overrides OGJK_DEVICES
  return 0;
}
int ogjk_release_cached_buffers(void) {
  // buffers cached by the calling thread on every device: host-path staging pools + scratch (kernels must be idle)
  if (t_pinned_flag) {
    cudaFreeHost(t_pinned_flag);
    t_pinned_flag = nullptr;
  }
  int cur = 0;
  OGJK_CK(cudaGetDevice(&cur));
  for (int d = 0; d < kMaxDevices; ++d) {
    DevicePool& p = t_pool[d];
    bool any = false;
    for (int k = 0; k < kSlotCount; ++k) any = any || p.ptr[k];
    if (!any) continue;
    OGJK_CK(cudaSetDevice(d));
    for (int k = 0; k < kSlotCount; ++k) {
      cudaFree(p.ptr[k]);
      p.ptr[k] = nullptr;
      p.cap[k] = 0;
    }
  }
  for (StreamScratch& x : t_sscratch) {
    OGJK_CK(cudaSetDevice(x.dev));
    cudaFree(x.ticket);
    x.ticket = nullptr;
    for (Scratch* sc : {&x.epa, &x.bp, &x.cr, &x.flag, &x.pack}) {
      cudaFree(sc->ptr);
      sc->ptr = nullptr;
      sc->ints = 0;
    }
  }
  OGJK_CK(cudaSetDevice(cur));
  return 0;
}
int ogjk_device_malloc(size_t bytes, void** d_ptr) {
  if (!d_ptr) return fail_msg("null argument");
  *d_ptr = nullptr;
  if (bytes == 0) return 0;
  OGJK_CK(cudaMalloc(d_ptr, bytes));
  OGJK_CK(cudaMemsetAsync(*d_ptr, 0, bytes, t_stream));  // zero-filled: results that EPA leaves unwritten are defined
  return 0;
}
int ogjk_device_free(void* d_ptr) {
  cudaFree(d_ptr);
  return 0;
}
int ogjk_memcpy_to_device(void* d_dst, const void* src, size_t bytes) {
  if (bytes == 0) return 0;
  OGJK_CK(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, t_stream));
  OGJK_CK(cudaStreamSynchronize(t_stream));
  return 0;
}
int ogjk_memcpy_from_device(void* dst, const void* d_src, size_t bytes) {
  if (bytes == 0) return 0;
  OGJK_CK(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, t_stream));
  OGJK_CK(cudaStreamSynchronize(t_stream));
  return 0;
}
const char* ogjk_last_kernel(void) { return t_last_kernel; }
long long ogjk_launch_count(int reset) {
  const long long v = t_launches;
  if (reset) t_launches = 0;
  return v;
}

#define OGJK_DEFINE_API(P, REAL)                                                                                       \
  int ogjk_##P##_compute_minimum_distance(int n, const void* bd1, const void* bd2, void* simplices, REAL* distances) { \
    return run_pairs_host_multi<REAL>(n, (const PolytopeT<REAL>*)bd1, (const PolytopeT<REAL>*)bd2,                           \
                                (SimplexT<REAL>*)simplices, distances, nullptr, nullptr, nullptr, kGjk);               \
  }                                                                                                                    \
  int ogjk_##P##_compute_collision_information(int n, const void* bd1, const void* bd2, void* simplices,              \
                                               REAL* distances, REAL* contact_normals) {                              \
    return run_pairs_host_multi<REAL>(n, (const PolytopeT<REAL>*)bd1, (const PolytopeT<REAL>*)bd2,                           \
                                (SimplexT<REAL>*)simplices, distances, contact_normals, nullptr, nullptr, kEpa);       \
  }                                                                                                                    \
  int ogjk_##P##_compute_gjk_epa(int n, const void* bd1, const void* bd2, void* simplices, REAL* distances,            \
                                 REAL* contact_normals) {                                                              \
    return run_pairs_host_multi<REAL>(n, (const PolytopeT<REAL>*)bd1, (const PolytopeT<REAL>*)bd2,                           \
                                (SimplexT<REAL>*)simplices, distances, contact_normals, nullptr, nullptr,              \
                                kGjk | kEpa);                                                                          \
  }                                                                                                                    \
  int ogjk_##P##_compute_collision_information_witness(int n, const void* bd1, const void* bd2, void* simplices,      \
                                                       REAL* distances, REAL* witness1, REAL* witness2,               \
                                                       REAL* contact_normals) {                                       \
    return run_pairs_host_multi<REAL>(n, (const PolytopeT<REAL>*)bd1, (const PolytopeT<REAL>*)bd2,                           \
                                (SimplexT<REAL>*)simplices, distances, contact_normals, witness1, witness2,            \
                                kGjk | kEpa);                                                                          \
  }                                                                                                                    \
  int ogjk_##P##_allocate_and_copy_device_arrays(int n, const void* bd1, const void* bd2, void** d_bd1,               \
                                                 void** d_bd2, REAL** d_coord1, REAL** d_coord2,                      \
                                                 void** d_simplices, REAL** d_distances) {                            \
    if (n <= 0) return 0;                                                                                              \
    Flattened<REAL> f1, f2;                                                                                            \
    if (int rc = flatten_upload<REAL>(n, (const PolytopeT<REAL>*)bd1, f1)) return rc;                                  \
    if (int rc = flatten_upload<REAL>(n, (const PolytopeT<REAL>*)bd2, f2)) {                                           \
      release(f1);                                                                                                     \
      return rc;                                                                                                       \
    }                                                                                                                  \
    *d_bd1 = f1.d_desc;                                                                                                \
    *d_bd2 = f2.d_desc;                                                                                                \
    *d_coord1 = f1.d_coord;                                                                                            \
    *d_coord2 = f2.d_coord;                                                                                            \
    cudaError_t e_ = cudaMalloc(d_simplices, (size_t)n * sizeof(SimplexT<REAL>));                                      \
    if (e_ == cudaSuccess) {                                                                                           \
      e_ = cudaMalloc((void**)d_distances, (size_t)n * sizeof(REAL));                                                  \
      if (e_ != cudaSuccess) cudaFree(*d_simplices);                                                                   \
    }                                                                                                                  \
    if (e_ == cudaSuccess) e_ = cudaMemsetAsync(*d_simplices, 0, (size_t)n * sizeof(SimplexT<REAL>), t_stream);        \
    if (e_ == cudaSuccess) e_ = cudaStreamSynchronize(t_stream);                                                       \
    if (e_ != cudaSuccess) {                                                                                           \
      release(f1);                                                                                                     \
      release(f2);                                                                                                     \
      *d_bd1 = *d_bd2 = *d_simplices = nullptr;                                                                        \
      *d_coord1 = *d_coord2 = *d_distances = nullptr;                                                                  \
      return fail("allocate_and_copy_device_arrays", e_);                                                              \
    }                                                                                                                  \
    /* remembered (after the last allocation succeeded) so that the *_device calls on these arrays can take the    */ \
    /* dense fast kernels; the layout is re-validated on the device at every use (pool_layout_holds)               */ \
    register_pool(f1.d_desc, f1.d_coord, f1.uniform_nv, n, (int)f1.max_nv);                                            \
    register_pool(f2.d_desc, f2.d_coord, f2.uniform_nv, n, (int)f2.max_nv);                                            \
    return 0;                                                                                                          \
  }                                                                                                                    \
  int ogjk_##P##_compute_minimum_distance_device(int n, const void* d_bd1, const void* d_bd2, void* d_simplices,      \
                                                 REAL* d_distances) {                                                 \
    if (n <= 0) return 0;                                                                                              \
    int nv = 0;                                                                                                        \
    PoolInfo p1, p2;                                                                                                   \
    const bool known = lookup_pool(d_bd1, &p1) && lookup_pool(d_bd2, &p2) && p1.count >= n && p2.count >= n;           \
    if (known && p1.nv > 0 && p2.nv > 0) { /* arrays uploaded by allocate_and_copy_device_arrays, uniform + dense */   \
      bool ok = false;                                                                                                 \
      if (int rc = pool_layouts_hold<REAL>(d_bd1, p1, d_bd2, &p2, n, &ok)) return rc;                                  \
      if (ok) {                                                                                                        \
        const int fast = launch_gjk_uniform<REAL>(n, p1.nv, (const REAL*)p1.coords, p2.nv, (const REAL*)p2.coords,     \
                                                  (SimplexT<REAL>*)d_simplices, d_distances);                          \
        if (fast <= 0) return fast;                                                                                    \
      }                                                                                                                \
    }                                                                                                                  \
    if (known) nv = p1.max_nv;                                                                                         \
    else if (int rc = peek_numpoints<REAL>((const PolytopeT<REAL>*)d_bd1, &nv)) return rc;                             \
    DescSource<REAL> src{(const PolytopeT<REAL>*)d_bd1, (const PolytopeT<REAL>*)d_bd2};                                \
    return launch_gjk_generic<REAL>(src, n, nv, (SimplexT<REAL>*)d_simplices, d_distances);                            \
  }                                                                                                                    \
  int ogjk_##P##_compute_epa_device(int n, const void* d_bd1, const void* d_bd2, void* d_simplices,                   \
                                    REAL* d_distances, REAL* d_contact_normals) {                                     \
    if (n <= 0) return 0;                                                                                              \
    DescSource<REAL> src{(const PolytopeT<REAL>*)d_bd1, (const PolytopeT<REAL>*)d_bd2};                                \
    /* vertex-count hint: selects lanes per pair only (any count is handled correctly by every kernel) */             \
    PoolInfo p1, p2;                                                                                                   \
    int nv = 0;                                                                                                        \
    if (lookup_pool(d_bd1, &p1) && lookup_pool(d_bd2, &p2)) nv = p1.max_nv > p2.max_nv ? p1.max_nv : p2.max_nv;        \
    else if (int rc = peek_numpoints<REAL>((const PolytopeT<REAL>*)d_bd1, &nv)) return rc;                             \
    return launch_epa<REAL>(src, n, nv, (SimplexT<REAL>*)d_simplices, d_distances, d_contact_normals);                 \
  }                                                                                                                    \
  int ogjk_##P##_copy_results_from_device(int n, const void* d_simplices, const REAL* d_distances, void* simplices,   \
                                          REAL* distances) {                                                          \
    if (n <= 0) return 0;                                                                                              \
    OGJK_CK(cudaMemcpyAsync(distances, d_distances, (size_t)n * sizeof(REAL), cudaMemcpyDeviceToHost, t_stream));      \
    OGJK_CK(cudaMemcpyAsync(simplices, d_simplices, (size_t)n * sizeof(SimplexT<REAL>), cudaMemcpyDeviceToHost,        \
                            t_stream));                                                                                \
    OGJK_CK(cudaStreamSynchronize(t_stream));                                                                          \
    return 0;                                                                                                          \
  }                                                                                                                    \
  int ogjk_##P##_free_device_arrays(void* d_bd1, void* d_bd2, REAL* d_coord1, REAL* d_coord2, void* d_simplices,      \
                                    REAL* d_distances) {                                                              \
    unregister_pool(d_bd1);                                                                                            \
    unregister_pool(d_bd2);                                                                                            \
    cudaFree(d_bd1);                                                                                                   \
    cudaFree(d_bd2);                                                                                                   \
    cudaFree(d_coord1);                                                                                                \
    cudaFree(d_coord2);                                                                                                \
    cudaFree(d_simplices);                                                                                             \
    cudaFree(d_distances);                                                                                             \
    return 0;                                                                                                          \
  }                                                                                                                    \
  int ogjk_##P##_allocate_epa_device_arrays(int n, REAL** d_witness1, REAL** d_witness2,                              \
                                            REAL** d_contact_normals) {                                               \
    if (n <= 0) return 0;                                                                                              \
    OGJK_CK(cudaMalloc((void**)d_witness1, (size_t)n * 3 * sizeof(REAL)));                                             \
    OGJK_CK(cudaMalloc((void**)d_witness2, (size_t)n * 3 * sizeof(REAL)));                                             \
    if (d_contact_normals) OGJK_CK(cudaMalloc((void**)d_contact_normals, (size_t)n * 3 * sizeof(REAL)));               \
    return 0;                                                                                                          \
  }                                                                                                                    \
  int ogjk_##P##_copy_epa_results_from_device(int n, const REAL* d_witness1, const REAL* d_witness2,                  \
                                              const REAL* d_contact_normals, REAL* witness1, REAL* witness2,          \
                                              REAL* contact_normals) {                                                \
    if (n <= 0) return 0;                                                                                              \
    const size_t bytes = (size_t)n * 3 * sizeof(REAL);                                                                 \
    OGJK_CK(cudaMemcpyAsync(witness1, d_witness1, bytes, cudaMemcpyDeviceToHost, t_stream));                           \
    OGJK_CK(cudaMemcpyAsync(witness2, d_witness2, bytes, cudaMemcpyDeviceToHost, t_stream));                           \
    if (contact_normals && d_contact_normals)                                                                          \
      OGJK_CK(cudaMemcpyAsync(contact_normals, d_contact_normals, bytes, cudaMemcpyDeviceToHost, t_stream));           \
    OGJK_CK(cudaStreamSynchronize(t_stream));                                                                          \
    return 0;                                                                                                          \
  }                                                                                                                    \
  int ogjk_##P##_free_epa_device_arrays(REAL* d_witness1, REAL* d_witness2, REAL* d_contact_normals) {                \
    cudaFree(d_witness1);                                                                                              \
    cudaFree(d_witness2);                                                                                              \
    cudaFree(d_contact_normals);                                                                                       \
    return 0;                                                                                                          \
  }                                                                                                                    \
  int ogjk_##P##_allocate_indexed_device(int num_polytopes, int max_pairs, const void* polytopes,                     \
                                         void** d_polytopes, REAL** d_coords, void** d_pairs, void** d_simplices,     \
                                         REAL** d_distances, REAL** d_contact_normals) {                              \
    if (num_polytopes <= 0 || max_pairs <= 0) return 0;                                                                \
    Flattened<REAL> pool;                                                                                              \
    if (int rc = flatten_upload<REAL>(num_polytopes, (const PolytopeT<REAL>*)polytopes, pool)) return rc;              \
    *d_polytopes = pool.d_desc;                                                                                        \
    *d_coords = pool.d_coord;                                                                                          \
    if (pool.uniform_nv > 0) register_pool(pool.d_desc, pool.d_coord, pool.uniform_nv, num_polytopes);                 \
    OGJK_CK(cudaMalloc(d_pairs, (size_t)max_pairs * sizeof(CollisionPair)));                                           \
    OGJK_CK(cudaMalloc(d_simplices, (size_t)max_pairs * sizeof(SimplexT<REAL>)));                                      \
    OGJK_CK(cudaMalloc((void**)d_distances, (size_t)max_pairs * sizeof(REAL)));                                        \
    if (d_contact_normals) OGJK_CK(cudaMalloc((void**)d_contact_normals, (size_t)max_pairs * 3 * sizeof(REAL)));       \
    return 0;                                                                                                          \
  }                                                                                                                    \
  int ogjk_##P##_free_indexed_device(void* d_polytopes, REAL* d_coords, void* d_pairs, void* d_simplices,             \
                                     REAL* d_distances, REAL* d_contact_normals) {                                    \
    unregister_pool(d_polytopes);                                                                                      \
    cudaFree(d_polytopes);                                                                                             \
    cudaFree(d_coords);                                                                                                \
    cudaFree(d_pairs);                                                                                                 \
    cudaFree(d_simplices);                                                                                             \
    cudaFree(d_distances);                                                                                             \
    cudaFree(d_contact_normals);                                                                                       \
    return 0;                                                                                                          \
  }                                                                                                                    \
  int ogjk_##P##_upload_pairs_device(int num_pairs, const void* pairs, void* d_pairs) {                               \
    if (num_pairs <= 0) return 0;                                                                                      \
    OGJK_CK(cudaMemcpyAsync(d_pairs, pairs, (size_t)num_pairs * sizeof(CollisionPair), cudaMemcpyHostToDevice,         \
                            t_stream));                                                                                \
    OGJK_CK(cudaStreamSynchronize(t_stream));                                                                          \
    return 0;                                                                                                          \
  }                                                                                                                    \
  int ogjk_##P##_compute_minimum_distance_indexed(int num_polytopes, int num_pairs, const void* polytopes,            \
                                                  const void* pairs, void* simplices, REAL* distances) {              \
    return run_indexed_host_multi<REAL>(num_polytopes, num_pairs, (const PolytopeT<REAL>*)polytopes,                         \
                                  (const CollisionPair*)pairs, (SimplexT<REAL>*)simplices, distances, nullptr, kGjk);  \
  }                                                                                                                    \
  int ogjk_##P##_compute_minimum_distance_indexed_device(int num_pairs, const void* d_polytopes,                      \
                                                         const void* d_pairs, void* d_simplices,                      \
                                                         REAL* d_distances) {                                         \
    if (num_pairs <= 0) return 0;                                                                                      \
    PoolInfo info;                                                                                                     \
    if (num_pairs >= 32768 && lookup_pool(d_polytopes, &info)) {                                                       \
      bool ok = false;                                                                                                 \
      if (int rc = pool_layout_holds<REAL>(d_polytopes, info, info.count, &ok)) return rc;                             \
      if (ok) {                                                                                                        \
        const int fast = launch_indexed_uniform<REAL>(num_pairs, info, (const CollisionPair*)d_pairs,                  \
                                                      (const PolytopeT<REAL>*)d_polytopes,                             \
                                                      (SimplexT<REAL>*)d_simplices, d_distances, nullptr, kGjkStage);  \
        if (fast <= 0) return fast;                                                                                    \
      }                                                                                                                \
    }                                                                                                                  \
    int nv = 0;                                                                                                        \
    if (int rc = peek_numpoints<REAL>((const PolytopeT<REAL>*)d_polytopes, &nv)) return rc;                            \
    IndexedSource<REAL> src{(const PolytopeT<REAL>*)d_polytopes, (const CollisionPair*)d_pairs};                       \
    return launch_gjk_generic<REAL>(src, num_pairs, nv, (SimplexT<REAL>*)d_simplices, d_distances);                    \
  }                                                                                                                    \
  int ogjk_##P##_gjk_epa_indexed_device(int num_pairs, const void* d_polytopes, const void* d_pairs,                  \
                                        void* d_simplices, REAL* d_distances, REAL* d_contact_normals) {              \
    if (num_pairs <= 0) return 0;                                                                                      \
    if (!d_contact_normals) return fail_msg("contact_normals must not be NULL on the device path");                    \
    PoolInfo info;                                                                                                     \
    if (num_pairs >= 32768 && lookup_pool(d_polytopes, &info)) {                                                       \
      bool ok = false;                                                                                                 \
      if (int rc = pool_layout_holds<REAL>(d_polytopes, info, info.count, &ok)) return rc;                             \
      if (ok) {                                                                                                        \
        const int fast = launch_indexed_uniform<REAL>(num_pairs, info, (const CollisionPair*)d_pairs,                  \
                                                      (const PolytopeT<REAL>*)d_polytopes,                             \
                                                      (SimplexT<REAL>*)d_simplices, d_distances, d_contact_normals,    \
                                                      kGjkStage | kEpaStage);                                          \
        if (fast <= 0) return fast;                                                                                    \
      }                                                                                                                \
    }                                                                                                                  \
    int nv = 0;                                                                                                        \
    if (int rc = peek_numpoints<REAL>((const PolytopeT<REAL>*)d_polytopes, &nv)) return rc;                            \
    IndexedSource<REAL> src{(const PolytopeT<REAL>*)d_polytopes, (const CollisionPair*)d_pairs};                       \
    const bool sync_saved = t_sync;                                                                                    \
    t_sync = false;                                                                                                    \
    int rc = stage_mark(0);                                                                                            \
    if (!rc) rc = launch_gjk_generic<REAL>(src, num_pairs, nv, (SimplexT<REAL>*)d_simplices, d_distances);             \
    t_sync = sync_saved;                                                                                               \
    if (!rc) rc = stage_mark(1);                                                                                       \
    if (!rc) rc = launch_epa<REAL>(src, num_pairs, nv, (SimplexT<REAL>*)d_simplices, d_distances, d_contact_normals);  \
    if (!rc) rc = stage_mark(2);                                                                                       \
    return rc;                                                                                                         \
  }                                                                                                                    \
  int ogjk_##P##_compute_epa_indexed_device(int num_pairs, const void* d_polytopes, const void* d_pairs,              \
                                            void* d_simplices, REAL* d_distances, REAL* d_contact_normals) {          \
    if (num_pairs <= 0) return 0;                                                                                      \
    IndexedSource<REAL> src{(const PolytopeT<REAL>*)d_polytopes, (const CollisionPair*)d_pairs};                       \
    PoolInfo info;                                                                                                     \
    int nv = 0;                                                                                                        \
    if (lookup_pool(d_polytopes, &info)) nv = info.max_nv;                                                             \
    else if (int rc = peek_numpoints<REAL>((const PolytopeT<REAL>*)d_polytopes, &nv)) return rc;                       \
    return launch_epa<REAL>(src, num_pairs, nv, (SimplexT<REAL>*)d_simplices, d_distances, d_contact_normals);         \
  }                                                                                                                    \
  int ogjk_##P##_compute_epa_indexed(int num_polytopes, int num_pairs, const void* polytopes, const void* pairs,      \
                                     void* simplices, REAL* distances, REAL* contact_normals) {                       \
    if (num_pairs <= 0) return 0;                                                                                      \
    return run_indexed_host_multi<REAL>(num_polytopes, num_pairs, (const PolytopeT<REAL>*)polytopes,                         \
                                  (const CollisionPair*)pairs, (SimplexT<REAL>*)simplices, distances,                  \
                                  contact_normals, kEpa);                                                              \
  }                                                                                                                    \
  int ogjk_##P##_compute_gjk_epa_indexed(int num_polytopes, int num_pairs, const void* polytopes,                     \
                                         const void* pairs, void* simplices, REAL* distances,                         \
                                         REAL* contact_normals) {                                                     \
    if (num_pairs <= 0) return 0;                                                                                      \
    return run_indexed_host_multi<REAL>(num_polytopes, num_pairs, (const PolytopeT<REAL>*)polytopes,                         \
                                  (const CollisionPair*)pairs, (SimplexT<REAL>*)simplices, distances,                  \
                                  contact_normals, kGjk | kEpa);                                                       \
  }                                                                                                                    \
  int ogjk_##P##_init_polytopes_device(void* d_polytopes, REAL* d_verts_world, const int* d_vert_offsets,             \
                                       const int* d_vert_counts, int uniform_count, int num_submeshes) {              \
    if (num_submeshes <= 0) return 0;                                                                                  \
    if (!d_polytopes || !d_verts_world) return fail_msg("null argument");                                              \
    const bool uniform = !d_vert_offsets || !d_vert_counts;                                                            \
    if (uniform && uniform_count < 1) return fail_msg("vertex layout missing");                                        \
    init_polytopes_kernel<REAL><<<(unsigned)((num_submeshes + 255) / 256), 256, 0, t_stream>>>(                        \
        (PolytopeT<REAL>*)d_polytopes, d_verts_world, uniform ? nullptr : d_vert_offsets,                              \
        uniform ? nullptr : d_vert_counts, uniform_count, num_submeshes);                                              \
    unregister_pool(d_polytopes);                                                                                      \
    if (uniform && uniform_count % 4 == 0 && ((uintptr_t)d_verts_world & 15u) == 0)                                    \
      register_pool(d_polytopes, d_verts_world, uniform_count, num_submeshes);                                         \
    return finish_launch("init_polytopes");                                                                            \
  }                                                                                                                    \
  int ogjk_##P##_contact_response_device(int num_pairs, const void* d_pairs, const REAL* d_distances,                  \
                                         const void* d_simplices, const REAL* d_contact_normals,                       \
                                         const int* d_sub_mesh_body, int num_objects, float* d_positions,              \
                                         const float* d_vel_ping, float* d_vel_pong, const float* d_ang_ping,          \
                                         float* d_ang_pong, const float* d_quats, const float* d_inv_inertia,          \
                                         const float* params) {                                                        \
    if (!params) return fail_msg("null argument");                                                                     \
    const ContactParams prm{params[0], params[1], params[2], params[3]};                                               \
    return contact_response<REAL>(num_pairs, (const CollisionPair*)d_pairs, d_distances,                               \
                                  (const SimplexT<REAL>*)d_simplices, d_contact_normals, d_sub_mesh_body, num_objects, \
                                  (float4*)d_positions, (const float4*)d_vel_ping, (float4*)d_vel_pong,                \
                                  (const float4*)d_ang_ping, (float4*)d_ang_pong, (const float4*)d_quats,              \
                                  d_inv_inertia, prm);                                                                 \
  }                                                                                                                    \
  int ogjk_##P##_gjk_uniform_device(int n, int nverts1, const REAL* d_coord1, int nverts2, const REAL* d_coord2,      \
                                    void* d_simplices, REAL* d_distances) {                                           \
    if (n <= 0) return 0;                                                                                              \
    if (nverts1 < 1 || nverts2 < 1) return fail_msg("polytope with no vertices");                                      \
    const int fast = launch_gjk_uniform<REAL>(n, nverts1, d_coord1, nverts2, d_coord2,                                 \
                                              (SimplexT<REAL>*)d_simplices, d_distances);                              \
    if (fast <= 0) return fast;                                                                                        \
    UniformSource<REAL> src{d_coord1, d_coord2, nverts1, nverts2};                                                     \
    return launch_gjk_generic<REAL>(src, n, (nverts1 + nverts2) / 2, (SimplexT<REAL>*)d_simplices, d_distances);       \
  }                                                                                                                    \
  int ogjk_##P##_gjk_epa_uniform_device(int n, int nverts1, const REAL* d_coord1, int nverts2,                        \
                                        const REAL* d_coord2, void* d_simplices, REAL* d_distances,                   \
                                        REAL* d_contact_normals) {                                                    \
    if (n <= 0) return 0;                                                                                              \
    if (nverts1 < 1 || nverts2 < 1) return fail_msg("polytope with no vertices");                                      \
    return launch_gjk_epa_uniform<REAL>(n, nverts1, d_coord1, nverts2, d_coord2, (SimplexT<REAL>*)d_simplices,         \
                                        d_distances, d_contact_normals);                                               \
  }                                                                                                                    \
  int ogjk_##P##_epa_uniform_device(int n, int nverts1, const REAL* d_coord1, int nverts2, const REAL* d_coord2,      \
                                    void* d_simplices, REAL* d_distances, REAL* d_contact_normals) {                  \
    if (n <= 0) return 0;                                                                                              \
    if (nverts1 < 1 || nverts2 < 1) return fail_msg("polytope with no vertices");                                      \
    UniformSource<REAL> src{d_coord1, d_coord2, nverts1, nverts2};                                                     \
    return launch_epa<REAL>(src, n, nverts1 > nverts2 ? nverts1 : nverts2, (SimplexT<REAL>*)d_simplices, d_distances,  \
                            d_contact_normals);                                                                        \
  }

OGJK_DEFINE_API(f32, float)
OGJK_DEFINE_API(f64, double)

}  // extern "C"

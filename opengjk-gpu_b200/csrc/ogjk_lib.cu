// ogjk_lib.cu -- the C ABI (include/opengjk_b200.h) over the sm_100a kernels.
//
// Host side of the drop-in boundary: the same allocate / flatten / launch / copy-back / free steps as the
// reference's wrappers (GJK/gpu/openGJK.cu:2787-3311), with checked CUDA calls, 64-bit size arithmetic (the
// reference accumulates coordinate counts in `int`, openGJK.cu:2910-2915), pinned staging and stream-ordered
// copies.  There is deliberately no CPU fallback: without a usable device every compute call fails.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/opengjk_b200.h"
#include "epa_kernel.cuh"
#include "gjk_generic.cuh"
#include "gjk_tables.h"
#include "ogjk_types.h"

using namespace ogjk;

namespace {

thread_local std::string t_err;
thread_local cudaStream_t t_stream = nullptr;
thread_local bool t_sync = true;
thread_local long long t_launches = 0;

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

int finish_launch(const char* what) {
  ++t_launches;
  OGJK_CK(cudaGetLastError());
  if (t_sync) {
    cudaError_t e = cudaStreamSynchronize(t_stream);
    if (e != cudaSuccess) return fail(what, e);
  }
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

template <typename T, typename Source>
int launch_epa(const Source& src, int n, SimplexT<T>* d_simplices, T* d_distances, T* d_normals) {
  if (!d_normals) return fail_msg("contact_normals must not be NULL on the device path");
  const int warps_per_block = EpaConfig<T>::kWarpsPerBlock;
  const unsigned grid = (unsigned)(((long long)n + warps_per_block - 1) / warps_per_block);
  epa_kernel<T, Source><<<grid, warps_per_block * 32, 0, t_stream>>>(src, d_simplices, d_distances, d_normals, n);
  return finish_launch("epa kernel");
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
};

template <typename T>
int flatten_upload(int n, const PolytopeT<T>* bd, Flattened<T>& out) {
  size_t total = 0;
  long long max_nv = 0;
  const size_t align = 16 / sizeof(T);
  for (int i = 0; i < n; ++i) {
    if (bd[i].numpoints < 1 || !bd[i].coord) return fail_msg("polytope with no vertices");
    total += ((size_t)bd[i].numpoints * 3 + align - 1) / align * align;
    if (bd[i].numpoints > max_nv) max_nv = bd[i].numpoints;
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

template <typename T>
int run_pairs_host(int n, const PolytopeT<T>* bd1, const PolytopeT<T>* bd2, SimplexT<T>* simplices, T* distances,
                   T* normals, T* witness1, T* witness2, int stages) {
  if (n <= 0) return 0;
  if (!bd1 || !bd2 || !simplices || !distances) return fail_msg("null argument");
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
      if ((stages & kEpa) && (rc = launch_epa<T>(src, n, d_simp, d_dist, d_nrm))) break;
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
  if (num_pairs <= 0 || num_polytopes <= 0) return 0;
  if (!polytopes || !pairs || !simplices || !distances) return fail_msg("null argument");
  Flattened<T> pool;
  CollisionPair* d_pairs = nullptr;
  SimplexT<T>* d_simp = nullptr;
  T* d_dist = nullptr;
  T* d_nrm = nullptr;
  int rc = 0;
  cudaError_t e;
  const size_t np = (size_t)num_pairs;
  do {
    for (size_t i = 0; i < np; ++i)
      if (pairs[i].idx1 < 0 || pairs[i].idx1 >= num_polytopes || pairs[i].idx2 < 0 || pairs[i].idx2 >= num_polytopes) {
        rc = fail_msg("pair index out of range");
        break;
      }
    if (rc) break;
    if ((rc = flatten_upload(num_polytopes, polytopes, pool))) break;
    if ((e = cudaMalloc(&d_pairs, np * sizeof(CollisionPair))) != cudaSuccess) { rc = fail("cudaMalloc(pairs)", e); break; }
    if ((e = cudaMalloc(&d_simp, np * sizeof(SimplexT<T>))) != cudaSuccess) { rc = fail("cudaMalloc(simplices)", e); break; }
    if ((e = cudaMalloc(&d_dist, np * sizeof(T))) != cudaSuccess) { rc = fail("cudaMalloc(distances)", e); break; }
    if ((e = cudaMemcpyAsync(d_pairs, pairs, np * sizeof(CollisionPair), cudaMemcpyHostToDevice, t_stream)) != cudaSuccess) { rc = fail("H2D pairs", e); break; }
    if (stages & kEpa) {
      if ((e = cudaMalloc(&d_nrm, np * 3 * sizeof(T))) != cudaSuccess) { rc = fail("cudaMalloc(normals)", e); break; }
      if ((e = cudaMemsetAsync(d_nrm, 0, np * 3 * sizeof(T), t_stream)) != cudaSuccess) { rc = fail("memset", e); break; }
    }
    if (stages & kGjk) {
      if ((e = cudaMemsetAsync(d_simp, 0, np * sizeof(SimplexT<T>), t_stream)) != cudaSuccess) { rc = fail("memset", e); break; }
    } else {
      if ((e = cudaMemcpyAsync(d_simp, simplices, np * sizeof(SimplexT<T>), cudaMemcpyHostToDevice, t_stream)) != cudaSuccess) { rc = fail("H2D simplices", e); break; }
      if ((e = cudaMemcpyAsync(d_dist, distances, np * sizeof(T), cudaMemcpyHostToDevice, t_stream)) != cudaSuccess) { rc = fail("H2D distances", e); break; }
    }
    {
      SyncOverride nosync;
      IndexedSource<T> src{pool.d_desc, d_pairs};
      if ((stages & kGjk) && (rc = launch_gjk_generic<T>(src, num_pairs, (int)pool.max_nv, d_simp, d_dist))) break;
      if ((stages & kEpa) && (rc = launch_epa<T>(src, num_pairs, d_simp, d_dist, d_nrm))) break;
    }
    if ((e = cudaMemcpyAsync(simplices, d_simp, np * sizeof(SimplexT<T>), cudaMemcpyDeviceToHost, t_stream)) != cudaSuccess) { rc = fail("D2H simplices", e); break; }
    if ((e = cudaMemcpyAsync(distances, d_dist, np * sizeof(T), cudaMemcpyDeviceToHost, t_stream)) != cudaSuccess) { rc = fail("D2H distances", e); break; }
    if ((stages & kEpa) && normals) {
      if ((e = cudaMemcpyAsync(normals, d_nrm, np * 3 * sizeof(T), cudaMemcpyDeviceToHost, t_stream)) != cudaSuccess) { rc = fail("D2H normals", e); break; }
    }
    if ((e = cudaStreamSynchronize(t_stream)) != cudaSuccess) { rc = fail("sync", e); break; }
  } while (0);
  release(pool);
  cudaFree(d_pairs);
  cudaFree(d_simp);
  cudaFree(d_dist);
  cudaFree(d_nrm);
  return rc;
}

}  // namespace

// =======================================================================================================
extern "C" {

const char* ogjk_last_error(void) { return t_err.c_str(); }
const char* ogjk_version(void) { return "opengjk-b200 0.1 (sm_100a)"; }
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
long long ogjk_launch_count(int reset) {
  const long long v = t_launches;
  if (reset) t_launches = 0;
  return v;
}

#define OGJK_DEFINE_API(P, REAL)                                                                                       \
  int ogjk_##P##_compute_minimum_distance(int n, const void* bd1, const void* bd2, void* simplices, REAL* distances) { \
    return run_pairs_host<REAL>(n, (const PolytopeT<REAL>*)bd1, (const PolytopeT<REAL>*)bd2,                           \
                                (SimplexT<REAL>*)simplices, distances, nullptr, nullptr, nullptr, kGjk);               \
  }                                                                                                                    \
  int ogjk_##P##_compute_collision_information(int n, const void* bd1, const void* bd2, void* simplices,              \
                                               REAL* distances, REAL* contact_normals) {                              \
    return run_pairs_host<REAL>(n, (const PolytopeT<REAL>*)bd1, (const PolytopeT<REAL>*)bd2,                           \
                                (SimplexT<REAL>*)simplices, distances, contact_normals, nullptr, nullptr, kEpa);       \
  }                                                                                                                    \
  int ogjk_##P##_compute_gjk_epa(int n, const void* bd1, const void* bd2, void* simplices, REAL* distances,            \
                                 REAL* contact_normals) {                                                              \
    return run_pairs_host<REAL>(n, (const PolytopeT<REAL>*)bd1, (const PolytopeT<REAL>*)bd2,                           \
                                (SimplexT<REAL>*)simplices, distances, contact_normals, nullptr, nullptr,              \
                                kGjk | kEpa);                                                                          \
  }                                                                                                                    \
  int ogjk_##P##_compute_collision_information_witness(int n, const void* bd1, const void* bd2, void* simplices,      \
                                                       REAL* distances, REAL* witness1, REAL* witness2,               \
                                                       REAL* contact_normals) {                                       \
    return run_pairs_host<REAL>(n, (const PolytopeT<REAL>*)bd1, (const PolytopeT<REAL>*)bd2,                           \
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
    OGJK_CK(cudaMalloc(d_simplices, (size_t)n * sizeof(SimplexT<REAL>)));                                              \
    OGJK_CK(cudaMalloc((void**)d_distances, (size_t)n * sizeof(REAL)));                                                \
    OGJK_CK(cudaMemsetAsync(*d_simplices, 0, (size_t)n * sizeof(SimplexT<REAL>), t_stream));                           \
    OGJK_CK(cudaStreamSynchronize(t_stream));                                                                          \
    return 0;                                                                                                          \
  }                                                                                                                    \
  int ogjk_##P##_compute_minimum_distance_device(int n, const void* d_bd1, const void* d_bd2, void* d_simplices,      \
                                                 REAL* d_distances) {                                                 \
    if (n <= 0) return 0;                                                                                              \
    int nv = 0;                                                                                                        \
    if (int rc = peek_numpoints<REAL>((const PolytopeT<REAL>*)d_bd1, &nv)) return rc;                                  \
    DescSource<REAL> src{(const PolytopeT<REAL>*)d_bd1, (const PolytopeT<REAL>*)d_bd2};                                \
    return launch_gjk_generic<REAL>(src, n, nv, (SimplexT<REAL>*)d_simplices, d_distances);                            \
  }                                                                                                                    \
  int ogjk_##P##_compute_epa_device(int n, const void* d_bd1, const void* d_bd2, void* d_simplices,                   \
                                    REAL* d_distances, REAL* d_contact_normals) {                                     \
    if (n <= 0) return 0;                                                                                              \
    DescSource<REAL> src{(const PolytopeT<REAL>*)d_bd1, (const PolytopeT<REAL>*)d_bd2};                                \
    return launch_epa<REAL>(src, n, (SimplexT<REAL>*)d_simplices, d_distances, d_contact_normals);                     \
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
    OGJK_CK(cudaMalloc(d_pairs, (size_t)max_pairs * sizeof(CollisionPair)));                                           \
    OGJK_CK(cudaMalloc(d_simplices, (size_t)max_pairs * sizeof(SimplexT<REAL>)));                                      \
    OGJK_CK(cudaMalloc((void**)d_distances, (size_t)max_pairs * sizeof(REAL)));                                        \
    if (d_contact_normals) OGJK_CK(cudaMalloc((void**)d_contact_normals, (size_t)max_pairs * 3 * sizeof(REAL)));       \
    return 0;                                                                                                          \
  }                                                                                                                    \
  int ogjk_##P##_free_indexed_device(void* d_polytopes, REAL* d_coords, void* d_pairs, void* d_simplices,             \
                                     REAL* d_distances, REAL* d_contact_normals) {                                    \
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
    return run_indexed_host<REAL>(num_polytopes, num_pairs, (const PolytopeT<REAL>*)polytopes,                         \
                                  (const CollisionPair*)pairs, (SimplexT<REAL>*)simplices, distances, nullptr, kGjk);  \
  }                                                                                                                    \
  int ogjk_##P##_compute_minimum_distance_indexed_device(int num_pairs, const void* d_polytopes,                      \
                                                         const void* d_pairs, void* d_simplices,                      \
                                                         REAL* d_distances) {                                         \
    if (num_pairs <= 0) return 0;                                                                                      \
    int nv = 0;                                                                                                        \
    if (int rc = peek_numpoints<REAL>((const PolytopeT<REAL>*)d_polytopes, &nv)) return rc;                            \
    IndexedSource<REAL> src{(const PolytopeT<REAL>*)d_polytopes, (const CollisionPair*)d_pairs};                       \
    return launch_gjk_generic<REAL>(src, num_pairs, nv, (SimplexT<REAL>*)d_simplices, d_distances);                    \
  }                                                                                                                    \
  int ogjk_##P##_compute_epa_indexed_device(int num_pairs, const void* d_polytopes, const void* d_pairs,              \
                                            void* d_simplices, REAL* d_distances, REAL* d_contact_normals) {          \
    if (num_pairs <= 0) return 0;                                                                                      \
    IndexedSource<REAL> src{(const PolytopeT<REAL>*)d_polytopes, (const CollisionPair*)d_pairs};                       \
    return launch_epa<REAL>(src, num_pairs, (SimplexT<REAL>*)d_simplices, d_distances, d_contact_normals);             \
  }                                                                                                                    \
  int ogjk_##P##_compute_epa_indexed(int num_polytopes, int num_pairs, const void* polytopes, const void* pairs,      \
                                     void* simplices, REAL* distances, REAL* contact_normals) {                       \
    if (num_pairs <= 0) return 0;                                                                                      \
    return run_indexed_host<REAL>(num_polytopes, num_pairs, (const PolytopeT<REAL>*)polytopes,                         \
                                  (const CollisionPair*)pairs, (SimplexT<REAL>*)simplices, distances,                  \
                                  contact_normals, kEpa);                                                              \
  }                                                                                                                    \
  int ogjk_##P##_compute_gjk_epa_indexed(int num_polytopes, int num_pairs, const void* polytopes,                     \
                                         const void* pairs, void* simplices, REAL* distances,                         \
                                         REAL* contact_normals) {                                                     \
    if (num_pairs <= 0) return 0;                                                                                      \
    return run_indexed_host<REAL>(num_polytopes, num_pairs, (const PolytopeT<REAL>*)polytopes,                         \
                                  (const CollisionPair*)pairs, (SimplexT<REAL>*)simplices, distances,                  \
                                  contact_normals, kGjk | kEpa);                                                       \
  }                                                                                                                    \
  int ogjk_##P##_gjk_uniform_device(int n, int nverts1, const REAL* d_coord1, int nverts2, const REAL* d_coord2,      \
                                    void* d_simplices, REAL* d_distances) {                                           \
    if (n <= 0) return 0;                                                                                              \
    if (nverts1 < 1 || nverts2 < 1) return fail_msg("polytope with no vertices");                                      \
    UniformSource<REAL> src{d_coord1, d_coord2, nverts1, nverts2};                                                     \
    return launch_gjk_generic<REAL>(src, n, (nverts1 + nverts2) / 2, (SimplexT<REAL>*)d_simplices, d_distances);       \
  }                                                                                                                    \
  int ogjk_##P##_epa_uniform_device(int n, int nverts1, const REAL* d_coord1, int nverts2, const REAL* d_coord2,      \
                                    void* d_simplices, REAL* d_distances, REAL* d_contact_normals) {                  \
    if (n <= 0) return 0;                                                                                              \
    if (nverts1 < 1 || nverts2 < 1) return fail_msg("polytope with no vertices");                                      \
    UniformSource<REAL> src{d_coord1, d_coord2, nverts1, nverts2};                                                     \
    return launch_epa<REAL>(src, n, (SimplexT<REAL>*)d_simplices, d_distances, d_contact_normals);                     \
  }

OGJK_DEFINE_API(f32, float)
OGJK_DEFINE_API(f64, double)

}  // extern "C"

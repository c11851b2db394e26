/*
 * oracle/ref_gpu_driver.cu -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin batch driver around the UNMODIFIED reference GPU library /root/reference/GJK/gpu/openGJK.cu, which
 * oracle/build_ref_gpu.sh compiles where it lies (nvcc -O3 --fmad=false, the reference's own flags,
 * GJK/CMakeLists.txt:32) into oracle/_ref_gpu/libogjk_refgpu_f32.so.  Two uses:
 *   * the "kernel to beat": bench.py times the reference's own device functions on the same B200 and the same
 *     pairs as ours (cudaEvent pair around compute_minimum_distance_device / compute_epa_device -- the reference's
 *     own timing definition, examples/gpu/example.cu:42-45, 75-78);
 *   * a second oracle: tests diff reference-GPU against reference-CPU on the seeded sets.
 *
 * Flat input format as oracle/ref_driver.c: dense [n][nv][3] coordinates per side, or one pool + int pairs.
 * Every function returns 0, or a cudaError_t.
 */
#include <cuda_runtime.h>

#include <vector>

#include "GJK/gpu/openGJK.h"

namespace {
struct Ev {
  cudaEvent_t a, b, c;
  Ev() {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventCreate(&c);
  }
  ~Ev() {
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaEventDestroy(c);
  }
};
void fill(std::vector<gkPolytope>& bd, const gkFloat* c, long n, int nv) {
  bd.resize((size_t)n);
  for (long i = 0; i < n; ++i) {
    gkPolytope& p = bd[(size_t)i];
    p.numpoints = nv;
    p.coord = const_cast<gkFloat*>(c) + (size_t)i * nv * 3;
    p.s[0] = p.s[1] = p.s[2] = 0;
    p.s_idx = 0;
  }
}
}  // namespace

extern "C" {

int ogjk_refgpu_sizeof_real(void) { return (int)sizeof(gkFloat); }
int ogjk_refgpu_sizeof_simplex(void) { return (int)sizeof(gkSimplex); }

/* Dense batch: GJK (do_gjk) then EPA (do_epa) exactly as the reference's compute_gjk_epa (openGJK.cu:2854-2883),
 * `reps` timed repetitions of the two device calls on the uploaded arrays; ms[0] / ms[1] = mean GJK / EPA time. */
int ogjk_refgpu_gjk_epa(long n, const gkFloat* c1, int nv1, const gkFloat* c2, int nv2, gkSimplex* simplices,
                        gkFloat* distances, gkFloat* normals, int do_epa, int reps, float* ms) {
  if (n <= 0) return 0;
  std::vector<gkPolytope> bd1, bd2;
  fill(bd1, c1, n, nv1);
  fill(bd2, c2, n, nv2);
  gkPolytope *d_bd1 = nullptr, *d_bd2 = nullptr;
  gkFloat *d_c1 = nullptr, *d_c2 = nullptr, *d_dist = nullptr, *d_nrm = nullptr;
  gkSimplex* d_simp = nullptr;
  allocate_and_copy_device_arrays((int)n, bd1.data(), bd2.data(), &d_bd1, &d_bd2, &d_c1, &d_c2, &d_simp, &d_dist);
  if (do_epa) cudaMalloc((void**)&d_nrm, (size_t)n * 3 * sizeof(gkFloat));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  Ev ev;
  float tg = 0, te = 0;
  if (reps < 1) reps = 1;
  for (int r = 0; r < reps; ++r) {
    cudaMemset(d_simp, 0, (size_t)n * sizeof(gkSimplex));
    if (d_nrm) cudaMemset(d_nrm, 0, (size_t)n * 3 * sizeof(gkFloat));
    cudaEventRecord(ev.a);
    compute_minimum_distance_device((int)n, d_bd1, d_bd2, d_simp, d_dist);
    cudaEventRecord(ev.b);
    if (do_epa) compute_epa_device((int)n, d_bd1, d_bd2, d_simp, d_dist, d_nrm);
    cudaEventRecord(ev.c);
    cudaEventSynchronize(ev.c);
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, ev.a, ev.b);
    cudaEventElapsedTime(&b, ev.b, ev.c);
    tg += a;
    te += b;
  }
  if (ms) {
    ms[0] = tg / reps;
    ms[1] = te / reps;
  }
  copy_results_from_device((int)n, d_simp, d_dist, simplices, distances);
  if (do_epa && normals) cudaMemcpy(normals, d_nrm, (size_t)n * 3 * sizeof(gkFloat), cudaMemcpyDeviceToHost);
  free_device_arrays(d_bd1, d_bd2, d_c1, d_c2, d_simp, d_dist);
  cudaFree(d_nrm);
  return (int)cudaGetLastError();
}

/* Indexed batch over one pool of uniform polytopes (the visualiser's per-frame call sequence,
 * visualization/integrate_final_gjk.cu:1028-1036). */
int ogjk_refgpu_gjk_epa_indexed(int npoly, int nv, const gkFloat* pool, long npairs, const int* pairs,
                                gkSimplex* simplices, gkFloat* distances, gkFloat* normals, int do_epa, int reps,
                                float* ms) {
  if (npairs <= 0 || npoly <= 0) return 0;
  std::vector<gkPolytope> bd;
  fill(bd, pool, npoly, nv);
  gkPolytope* d_poly = nullptr;
  gkFloat *d_coords = nullptr, *d_dist = nullptr, *d_nrm = nullptr;
  gkCollisionPair* d_pairs = nullptr;
  gkSimplex* d_simp = nullptr;
  allocate_indexed_device(npoly, (int)npairs, bd.data(), &d_poly, &d_coords, &d_pairs, &d_simp, &d_dist, &d_nrm);
  upload_pairs_device((int)npairs, reinterpret_cast<const gkCollisionPair*>(pairs), d_pairs);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  Ev ev;
  float tg = 0, te = 0;
  if (reps < 1) reps = 1;
  for (int r = 0; r < reps; ++r) {
    cudaMemset(d_simp, 0, (size_t)npairs * sizeof(gkSimplex));
    cudaMemset(d_nrm, 0, (size_t)npairs * 3 * sizeof(gkFloat));
    cudaEventRecord(ev.a);
    compute_minimum_distance_indexed_device((int)npairs, d_poly, d_pairs, d_simp, d_dist);
    cudaEventRecord(ev.b);
    if (do_epa) compute_epa_indexed_device((int)npairs, d_poly, d_pairs, d_simp, d_dist, d_nrm);
    cudaEventRecord(ev.c);
    cudaEventSynchronize(ev.c);
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, ev.a, ev.b);
    cudaEventElapsedTime(&b, ev.b, ev.c);
    tg += a;
    te += b;
  }
  if (ms) {
    ms[0] = tg / reps;
    ms[1] = te / reps;
  }
  cudaMemcpy(simplices, d_simp, (size_t)npairs * sizeof(gkSimplex), cudaMemcpyDeviceToHost);
  cudaMemcpy(distances, d_dist, (size_t)npairs * sizeof(gkFloat), cudaMemcpyDeviceToHost);
  if (do_epa && normals) cudaMemcpy(normals, d_nrm, (size_t)npairs * 3 * sizeof(gkFloat), cudaMemcpyDeviceToHost);
  free_indexed_device(d_poly, d_coords, d_pairs, d_simp, d_dist, d_nrm);
  return (int)cudaGetLastError();
}

}  // extern "C"

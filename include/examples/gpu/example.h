/*
 * examples/gpu/example.h -- drop-in for the reference's namespace API (reference examples/gpu/example.h:7-96,
 * implemented at examples/gpu/example.cu:13-119): GJK::GPU::computeDistances / computeEPA / computeGJKAndEPA with
 * the same signatures, plus the README spelling computeCollisionInformation(..., witness1, witness2,
 * contact_normals = nullptr) (reference README.md:38-47) that the reference documents but never defines.
 * As in the reference (examples/gpu/example.cu:42-45, 75-78, 110-112), timer() brackets only the *_device launches --
 * not allocation, upload, download or free -- for all three entry points.
 * Multi-GPU: when more than one device has been selected (ogjk_set_devices() or OGJK_DEVICES in the environment) the
 * three entry points hand the host arrays to the library's host-pointer calls, which slice the pair range over the
 * devices; timer() then brackets the whole fanned-out call (the kernels of different devices have no common clock).
 */
#ifndef EXAMPLE_H
#define EXAMPLE_H

#include "../../GJK/gpu/openGJK.h"
#include "../common/timer.h"

namespace GJK {
namespace GPU {

inline GJK::Common::PerformanceTimer& timer() {
  static GJK::Common::PerformanceTimer t;
  return t;
}

inline void computeDistances(const int n, const gkPolytope* bd1, const gkPolytope* bd2, gkSimplex* simplices,
                             gkFloat* distances) {
  if (n <= 0) return;
  if (ogjk_selected_device_count() > 1) {
    timer().startGpuTimer();
    compute_minimum_distance(n, bd1, bd2, simplices, distances);
    timer().endGpuTimer();
    return;
  }
  gkPolytope *d_bd1 = nullptr, *d_bd2 = nullptr;
  gkFloat *d_coord1 = nullptr, *d_coord2 = nullptr, *d_distances = nullptr;
  gkSimplex* d_simplices = nullptr;
  allocate_and_copy_device_arrays(n, bd1, bd2, &d_bd1, &d_bd2, &d_coord1, &d_coord2, &d_simplices, &d_distances);
  timer().startGpuTimer();
  compute_minimum_distance_device(n, d_bd1, d_bd2, d_simplices, d_distances);
  timer().endGpuTimer();
  copy_results_from_device(n, d_simplices, d_distances, simplices, distances);
  free_device_arrays(d_bd1, d_bd2, d_coord1, d_coord2, d_simplices, d_distances);
}

namespace detail {
inline gkFloat* alloc_normals(const int n) {
  void* p = nullptr;
  ogjk_device_malloc((size_t)n * 3 * sizeof(gkFloat), &p);
  return static_cast<gkFloat*>(p);
}
}  // namespace detail

inline void computeEPA(const int n, const gkPolytope* bd1, const gkPolytope* bd2, gkSimplex* simplices,
                       gkFloat* distances, gkFloat* contact_normals) {
  if (n <= 0) return;
  if (ogjk_selected_device_count() > 1) {
    timer().startGpuTimer();
    computeCollisionInformation(n, bd1, bd2, simplices, distances, contact_normals);
    timer().endGpuTimer();
    return;
  }
  gkPolytope *d_bd1 = nullptr, *d_bd2 = nullptr;
  gkFloat *d_coord1 = nullptr, *d_coord2 = nullptr, *d_distances = nullptr;
  gkSimplex* d_simplices = nullptr;
  allocate_and_copy_device_arrays(n, bd1, bd2, &d_bd1, &d_bd2, &d_coord1, &d_coord2, &d_simplices, &d_distances);
  ogjk_memcpy_to_device(d_simplices, simplices, (size_t)n * sizeof(gkSimplex));
  ogjk_memcpy_to_device(d_distances, distances, (size_t)n * sizeof(gkFloat));
  gkFloat* d_contact_normals = detail::alloc_normals(n);
  timer().startGpuTimer();
  compute_epa_device(n, d_bd1, d_bd2, d_simplices, d_distances, d_contact_normals);
  timer().endGpuTimer();
  copy_results_from_device(n, d_simplices, d_distances, simplices, distances);
  ogjk_memcpy_from_device(contact_normals, d_contact_normals, (size_t)n * 3 * sizeof(gkFloat));
  free_device_arrays(d_bd1, d_bd2, d_coord1, d_coord2, d_simplices, d_distances);
  ogjk_device_free(d_contact_normals);
}

inline void computeGJKAndEPA(const int n, const gkPolytope* bd1, const gkPolytope* bd2, gkSimplex* simplices,
                             gkFloat* distances, gkFloat* contact_normals) {
  if (n <= 0) return;
  if (ogjk_selected_device_count() > 1) {
    timer().startGpuTimer();
    compute_gjk_epa(n, bd1, bd2, simplices, distances, contact_normals);
    timer().endGpuTimer();
    return;
  }
  gkPolytope *d_bd1 = nullptr, *d_bd2 = nullptr;
  gkFloat *d_coord1 = nullptr, *d_coord2 = nullptr, *d_distances = nullptr;
  gkSimplex* d_simplices = nullptr;
  allocate_and_copy_device_arrays(n, bd1, bd2, &d_bd1, &d_bd2, &d_coord1, &d_coord2, &d_simplices, &d_distances);
  gkFloat* d_contact_normals = detail::alloc_normals(n);
  timer().startGpuTimer();
  compute_minimum_distance_device(n, d_bd1, d_bd2, d_simplices, d_distances);
  compute_epa_device(n, d_bd1, d_bd2, d_simplices, d_distances, d_contact_normals);
  timer().endGpuTimer();
  copy_results_from_device(n, d_simplices, d_distances, simplices, distances);
  ogjk_memcpy_from_device(contact_normals, d_contact_normals, (size_t)n * 3 * sizeof(gkFloat));
  free_device_arrays(d_bd1, d_bd2, d_coord1, d_coord2, d_simplices, d_distances);
  ogjk_device_free(d_contact_normals);
}

/* README spelling: GJK + EPA with the witness points also returned as two n x 3 arrays */
inline void computeCollisionInformation(const int n, const gkPolytope* bd1, const gkPolytope* bd2,
                                        gkSimplex* simplices, gkFloat* distances, gkFloat* witness1,
                                        gkFloat* witness2, gkFloat* contact_normals = nullptr) {
  if (n <= 0) return;
  OGJK_API(compute_collision_information_witness)(n, bd1, bd2, simplices, distances, witness1, witness2,
                                                  contact_normals);
}

}  // namespace GPU
}  // namespace GJK

#endif /* EXAMPLE_H */

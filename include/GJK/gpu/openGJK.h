/*
 * GJK/gpu/openGJK.h -- drop-in replacement for the reference's GPU API header (reference GJK/gpu/openGJK.h:91-505).
 *
 * Every host function of the reference's library layer is provided with the identical signature, as an inline
 * wrapper over the C ABI in opengjk_b200.h (link with -lopengjk_b200; no nvcc needed to compile callers).  Like the
 * reference they return void; a failure leaves a message in ogjk_last_error().  The reference's four __global__
 * kernel declarations (openGJK.h:50-51, 68-69, 376-382, 428-435) are not re-exported: no caller in the reference
 * launches them directly and the kernels here have a different work decomposition.
 */
#ifndef OPENGJK_H__
#define OPENGJK_H__

#include "../common.h"
#include "../../opengjk_b200.h"

struct gkCollisionPair {
  int idx1; /* index of the first polytope in the pool  */
  int idx2; /* index of the second polytope in the pool */
};

/* ---- high level: host pointers in, host pointers out --------------------------------------------------------- */
inline void compute_minimum_distance(const int n, const gkPolytope* bd1, const gkPolytope* bd2, gkSimplex* simplices,
                                     gkFloat* distances) {
  OGJK_API(compute_minimum_distance)(n, bd1, bd2, simplices, distances);
}
inline void computeCollisionInformation(const int n, const gkPolytope* bd1, const gkPolytope* bd2,
                                        gkSimplex* simplices, gkFloat* distances, gkFloat* contact_normals) {
  OGJK_API(compute_collision_information)(n, bd1, bd2, simplices, distances, contact_normals);
}
inline void compute_gjk_epa(const int n, const gkPolytope* bd1, const gkPolytope* bd2, gkSimplex* simplices,
                            gkFloat* distances, gkFloat* contact_normals) {
  OGJK_API(compute_gjk_epa)(n, bd1, bd2, simplices, distances, contact_normals);
}

/* ---- mid level: explicit device memory ------------------------------------------------------------------------ */
inline void allocate_epa_device_arrays(const int n, gkFloat** d_witness1, gkFloat** d_witness2,
                                       gkFloat** d_contact_normals) {
  OGJK_API(allocate_epa_device_arrays)(n, d_witness1, d_witness2, d_contact_normals);
}
inline void copy_epa_results_from_device(const int n, const gkFloat* d_witness1, const gkFloat* d_witness2,
                                         const gkFloat* d_contact_normals, gkFloat* witness1, gkFloat* witness2,
                                         gkFloat* contact_normals) {
  OGJK_API(copy_epa_results_from_device)(n, d_witness1, d_witness2, d_contact_normals, witness1, witness2,
                                         contact_normals);
}
inline void free_epa_device_arrays(gkFloat* d_witness1, gkFloat* d_witness2, gkFloat* d_contact_normals) {
  OGJK_API(free_epa_device_arrays)(d_witness1, d_witness2, d_contact_normals);
}
inline void allocate_and_copy_device_arrays(const int n, const gkPolytope* bd1, const gkPolytope* bd2,
                                            gkPolytope** d_bd1, gkPolytope** d_bd2, gkFloat** d_coord1,
                                            gkFloat** d_coord2, gkSimplex** d_simplices, gkFloat** d_distances) {
  OGJK_API(allocate_and_copy_device_arrays)(n, bd1, bd2, (void**)d_bd1, (void**)d_bd2, d_coord1, d_coord2,
                                            (void**)d_simplices, d_distances);
}
inline void compute_minimum_distance_device(const int n, const gkPolytope* d_bd1, const gkPolytope* d_bd2,
                                            gkSimplex* d_simplices, gkFloat* d_distances) {
  OGJK_API(compute_minimum_distance_device)(n, d_bd1, d_bd2, d_simplices, d_distances);
}
inline void compute_epa_device(const int n, const gkPolytope* d_bd1, const gkPolytope* d_bd2, gkSimplex* d_simplices,
                               gkFloat* d_distances, gkFloat* d_contact_normals) {
  OGJK_API(compute_epa_device)(n, d_bd1, d_bd2, d_simplices, d_distances, d_contact_normals);
}
inline void copy_results_from_device(const int n, const gkSimplex* d_simplices, const gkFloat* d_distances,
                                     gkSimplex* simplices, gkFloat* distances) {
  OGJK_API(copy_results_from_device)(n, d_simplices, d_distances, simplices, distances);
}
inline void free_device_arrays(gkPolytope* d_bd1, gkPolytope* d_bd2, gkFloat* d_coord1, gkFloat* d_coord2,
                               gkSimplex* d_simplices, gkFloat* d_distances) {
  OGJK_API(free_device_arrays)(d_bd1, d_bd2, d_coord1, d_coord2, d_simplices, d_distances);
}

/* ---- indexed: one polytope pool + index pairs ------------------------------------------------------------------ */
inline void allocate_indexed_device(const int num_polytopes, const int max_pairs, const gkPolytope* polytopes,
                                    gkPolytope** d_polytopes, gkFloat** d_coords, gkCollisionPair** d_pairs,
                                    gkSimplex** d_simplices, gkFloat** d_distances, gkFloat** d_contact_normals) {
  OGJK_API(allocate_indexed_device)(num_polytopes, max_pairs, polytopes, (void**)d_polytopes, d_coords,
                                    (void**)d_pairs, (void**)d_simplices, d_distances, d_contact_normals);
}
inline void free_indexed_device(gkPolytope* d_polytopes, gkFloat* d_coords, gkCollisionPair* d_pairs,
                                gkSimplex* d_simplices, gkFloat* d_distances, gkFloat* d_contact_normals) {
  OGJK_API(free_indexed_device)(d_polytopes, d_coords, d_pairs, d_simplices, d_distances, d_contact_normals);
}
inline void upload_pairs_device(const int num_pairs, const gkCollisionPair* pairs, gkCollisionPair* d_pairs) {
  OGJK_API(upload_pairs_device)(num_pairs, pairs, d_pairs);
}
inline void compute_minimum_distance_indexed(const int num_polytopes, const int num_pairs,
                                             const gkPolytope* polytopes, const gkCollisionPair* pairs,
                                             gkSimplex* simplices, gkFloat* distances) {
  OGJK_API(compute_minimum_distance_indexed)(num_polytopes, num_pairs, polytopes, pairs, simplices, distances);
}
inline void compute_minimum_distance_indexed_device(const int num_pairs, const gkPolytope* d_polytopes,
                                                    const gkCollisionPair* d_pairs, gkSimplex* d_simplices,
                                                    gkFloat* d_distances) {
  OGJK_API(compute_minimum_distance_indexed_device)(num_pairs, d_polytopes, d_pairs, d_simplices, d_distances);
}
inline void compute_epa_indexed_device(const int num_pairs, const gkPolytope* d_polytopes,
                                       const gkCollisionPair* d_pairs, gkSimplex* d_simplices, gkFloat* d_distances,
                                       gkFloat* d_contact_normals) {
  OGJK_API(compute_epa_indexed_device)(num_pairs, d_polytopes, d_pairs, d_simplices, d_distances, d_contact_normals);
}
inline void compute_epa_indexed(const int num_polytopes, const int num_pairs, const gkPolytope* polytopes,
                                const gkCollisionPair* pairs, gkSimplex* simplices, gkFloat* distances,
                                gkFloat* contact_normals) {
  OGJK_API(compute_epa_indexed)(num_polytopes, num_pairs, polytopes, pairs, simplices, distances, contact_normals);
}
inline void compute_gjk_epa_indexed(const int num_polytopes, const int num_pairs, const gkPolytope* polytopes,
                                    const gkCollisionPair* pairs, gkSimplex* simplices, gkFloat* distances,
                                    gkFloat* contact_normals) {
  OGJK_API(compute_gjk_epa_indexed)(num_polytopes, num_pairs, polytopes, pairs, simplices, distances,
                                    contact_normals);
}

#endif /* OPENGJK_H__ */

/*
 * opengjk_b200.h -- C ABI of the B200-native batched GJK distance / EPA penetration engine.
 *
 * This is the drop-in boundary for the reference's GPU path: every function below replaces one host
 * function of the reference's library layer (GJK/gpu/openGJK.h, implemented at GJK/gpu/openGJK.cu:2787-3311)
 * and has the same arguments with the same meaning.  Differences, all of them additive:
 *   - plain C linkage, one symbol set per precision: ogjk_f32_* (reference built with USE_32BITS,
 *     GJK/common.h:44) and ogjk_f64_* (USE_32BITS commented out).  `OGJK_REAL` below is float / double;
 *     gkPolytope / gkSimplex / gkCollisionPair arrays are passed as `void*` and must have the reference's
 *     layouts for that precision (GJK/common.h:68-89; fp32 32/108 B, fp64 48/184 B; pair = 2 x int).
 *   - every function returns an int status (0 = ok) instead of void; the message of the last failure on the
 *     calling thread is available from ogjk_last_error().  The reference checks no CUDA call at all.
 *   - launches go to the stream selected with ogjk_set_stream() (default: the legacy default stream the
 *     reference uses) and, like the reference (openGJK.cu:2983, 3003, 3141, 3161), each *_device call ends with a
 *     device synchronisation unless ogjk_set_sync(0) was called.  Asynchronous calls of one thread may be in flight
 *     on several streams at once: the library's scratch (tickets, EPA queue, ...) is kept per (device, stream).
 *   - device descriptor arrays returned by allocate_* are remembered so that *_device calls on them can take the
 *     dense fast kernels; the remembered layout is re-validated against the live descriptors on the device at every
 *     call, so editing or re-pointing descriptors after upload is allowed (it selects the general kernels).
 * Semantics kept from the reference: n <= 0 is a silent no-op; simplices[i] on input and bd.s / bd.s_idx are
 * ignored by GJK; EPA writes the penetration depth as a NEGATIVE distance and leaves pairs with
 * distance > epsilon untouched except for the contact normal.
 *
 * The C++ spellings a user of the reference already calls (compute_minimum_distance(...),
 * GJK::GPU::computeDistances(...), ...) are inline forwarding wrappers in include/GJK/gpu/openGJK.h and
 * include/examples/gpu/example.h.
 */
#ifndef OPENGJK_B200_H__
#define OPENGJK_B200_H__

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- process-wide helpers ---------------------------------------------------------------------------- */
/* Message of the last failed call on this thread ("" if none).  Replaces nothing: the reference has no
 * error reporting. */
const char* ogjk_last_error(void);
/* Library version string, e.g. "opengjk-b200 0.1 (sm_100a)". */
const char* ogjk_version(void);
/* Number of CUDA devices visible, or -1 on error. */
int ogjk_device_count(void);
/* Select the CUDA device subsequent calls on this thread use (the reference uses the current device). */
int ogjk_set_device(int device);
/* Stream (a cudaStream_t cast to void*) used by subsequent launches/copies on this thread; NULL = default. */
int ogjk_set_stream(void* stream);
/* 1 (default): *_device calls synchronise the device before returning, as the reference does; 0: async. */
int ogjk_set_sync(int enabled);
/* Multi-GPU (the reference is single-device).  With more than one device selected, the
 * host-pointer entry points (compute_minimum_distance, compute_collision_information[_witness], compute_gjk_epa and
 * the three *_indexed host calls) split the pair range into one contiguous slice per device; a persistent host
 * thread per device runs the single-device path on its slice and copies the results straight into the caller's
 * arrays at the slice offset (indexed calls: the pool is replicated, the pair list sliced).  No collective, nothing
 * exchanged.  devices == NULL selects ordinals 0..count-1; count <= 1 restores single-device behaviour; a device
 * listed more than once gets that many slices, run back to back.  The same
 * selection can be made without touching the caller's code through the environment: OGJK_DEVICES=all | <count> |
 * <i,j,...>.  Process-wide.  The *_device entry points always use the calling thread's current device. */
int ogjk_set_devices(int count, const int* devices);
/* Number of devices the host-pointer entry points currently fan out over (0 or 1: single-device behaviour). */
int ogjk_selected_device_count(void);
/* Frees the device buffers the calling thread has cached (the host-pointer path keeps its staging buffers between
 * calls instead of cudaMalloc/cudaFree per call as the reference does, openGJK.cu:2889-2954, 3034-3048; scratch of the
 * EPA queue / broad phase / contact response).  No call of this thread may be in flight. */
int ogjk_release_cached_buffers(void);
/* Plain device-memory helpers so that a caller without the CUDA toolkit headers (the drop-in example.h) can follow
 * the reference's allocate -> upload -> launch -> download -> free sequence (examples/gpu/example.cu:86-119). */
int ogjk_device_malloc(size_t bytes, void** d_ptr); /* zero-filled */
int ogjk_device_free(void* d_ptr);
int ogjk_memcpy_to_device(void* d_dst, const void* src, size_t bytes);
int ogjk_memcpy_from_device(void* dst, const void* d_src, size_t bytes);
/* Number of kernels this library has launched on the calling thread since the last reset (bench accounting). */
long long ogjk_launch_count(int reset);
/* Name of the kernel family the calling thread launched last ("gjk slots (fp16 pre-scan) kernel", "epa kernel", ...):
 * lets a test or a benchmark state which kernel produced a result. */
const char* ogjk_last_kernel(void);
/* Uniform-grid broad phase on the device (the step before the hot path in the reference's caller:
 * visualization/integrate_final_gjk.cu:467-570 insert/count/generate kernels, :916-1002 sim_broad_phase).
 * d_pos_radius: num_objects x float4 = centre xyz + bounding radius.  An object lives in the cell
 * floor((p + boundary) / cell_size) clamped to [0, grid_size); object i is paired with every j > i in the 27
 * surrounding cells with |ci - cj|^2 < (ri + rj)^2 (fp32, unfused).  Pairs are written to d_pairs
 * (gkCollisionPair[max_pairs]) grouped by i ascending, writes beyond max_pairs are dropped; *num_pairs receives the
 * number of pairs found (compare with max_pairs to detect the clamp).  The list feeds
 * ogjk_*_compute_minimum_distance_indexed_device / ogjk_*_compute_epa_indexed_device directly. */
int ogjk_broadphase_pairs_device(int num_objects, const float* d_pos_radius, float cell_size, float boundary,
                                 int grid_size, void* d_pairs, int max_pairs, long long* num_pairs);
/* Local -> world vertex transform of the polytope pool (reference visualization/integrate_final_gjk.cu:304-332
 * transform_to_world_kernel): world = quat_rotate(quats[b], local * scales[b]) + positions[b] for every vertex of
 * sub-mesh sm, b = sub_mesh_body[sm].  positions / quats: float4 per body (xyz + unused w / xyzw); scales: 3 floats
 * per body; vertices: 3 floats each.  vert_offsets / vert_counts / sub_mesh_body may be NULL: then every sub-mesh has
 * uniform_count vertices at offset sm * uniform_count and belongs to body sm.  fp32, as in the reference. */
int ogjk_transform_to_world_device(int num_submeshes, const float* d_positions, const float* d_quats,
                                   const float* d_scales, const float* d_verts_local, float* d_verts_world,
                                   const int* d_vert_offsets, const int* d_vert_counts, const int* d_sub_mesh_body,
                                   int uniform_count);
/* Forget a pool that ogjk_*_init_polytopes_device registered for the slot-kernel fast path (call before freeing a
 * caller-owned descriptor array; ogjk_*_free_indexed_device does it for pools the library allocated). */
int ogjk_release_pool(const void* d_polytopes);
/* Per-stage device timing of the fused *_gjk_epa_uniform_device calls of this thread: while enabled every call
 * records CUDA events on the launching stream before GJK, between the stages and after EPA; ogjk_stage_times waits
 * for them, returns the summed GJK / EPA milliseconds and the number of calls, and resets the list. */
int ogjk_set_timing(int enabled);
int ogjk_stage_times(double* gjk_ms, double* epa_ms, int* calls);

#define OGJK_DECLARE_API(P, OGJK_REAL)                                                                          \
  /* ---- high level: host pointers, device memory handled internally -------------------------------- */     \
  /* reference: compute_minimum_distance, GJK/gpu/openGJK.h:91-97 (openGJK.cu:2791-2819) */                     \
  int ogjk_##P##_compute_minimum_distance(int n, const void* bd1, const void* bd2, void* simplices,            \
                                          OGJK_REAL* distances);                                                \
  /* reference: computeCollisionInformation, openGJK.h:112-119 (openGJK.cu:2821-2852); uploads the caller's   \
   * simplices AND distances (the reference forgets the latter, SURVEY.md section 7 quirk iv) */               \
  int ogjk_##P##_compute_collision_information(int n, const void* bd1, const void* bd2, void* simplices,       \
                                               OGJK_REAL* distances, OGJK_REAL* contact_normals);              \
  /* reference: compute_gjk_epa, openGJK.h:134-141 (openGJK.cu:2854-2883) */                                    \
  int ogjk_##P##_compute_gjk_epa(int n, const void* bd1, const void* bd2, void* simplices,                     \
                                 OGJK_REAL* distances, OGJK_REAL* contact_normals);                            \
  /* README spelling GJK::GPU::computeCollisionInformation(n, bd1, bd2, simplices, distances, witness1,        \
   * witness2, contact_normals) (reference README.md:38-47): GJK + EPA, witnesses also copied to n x 3 arrays;  \
   * witness1 / witness2 / contact_normals may be NULL */                                                       \
  int ogjk_##P##_compute_collision_information_witness(int n, const void* bd1, const void* bd2,                \
                                                       void* simplices, OGJK_REAL* distances,                  \
                                                       OGJK_REAL* witness1, OGJK_REAL* witness2,               \
                                                       OGJK_REAL* contact_normals);                            \
  /* ---- mid level: explicit device memory -------------------------------------------------------------- */ \
  /* reference: allocate_and_copy_device_arrays, openGJK.h:212-222 (openGJK.cu:2889-2954).  d_simplices is     \
   * zero-filled; descriptors in *d_bd1 / *d_bd2 point into *d_coord1 / *d_coord2 (flattened xyz). */          \
  int ogjk_##P##_allocate_and_copy_device_arrays(int n, const void* bd1, const void* bd2, void** d_bd1,        \
                                                 void** d_bd2, OGJK_REAL** d_coord1, OGJK_REAL** d_coord2,     \
                                                 void** d_simplices, OGJK_REAL** d_distances);                 \
  /* reference: compute_minimum_distance_device, openGJK.h:237-243 (openGJK.cu:2968-2984) */                    \
  int ogjk_##P##_compute_minimum_distance_device(int n, const void* d_bd1, const void* d_bd2,                  \
                                                 void* d_simplices, OGJK_REAL* d_distances);                   \
  /* reference: compute_epa_device, openGJK.h:255-262 (openGJK.cu:2986-3004) */                                 \
  int ogjk_##P##_compute_epa_device(int n, const void* d_bd1, const void* d_bd2, void* d_simplices,            \
                                    OGJK_REAL* d_distances, OGJK_REAL* d_contact_normals);                     \
  /* reference: copy_results_from_device, openGJK.h:273-279 (openGJK.cu:3006-3015) */                           \
  int ogjk_##P##_copy_results_from_device(int n, const void* d_simplices, const OGJK_REAL* d_distances,        \
                                          void* simplices, OGJK_REAL* distances);                              \
  /* reference: free_device_arrays, openGJK.h:292-299 (openGJK.cu:3034-3048) */                                 \
  int ogjk_##P##_free_device_arrays(void* d_bd1, void* d_bd2, OGJK_REAL* d_coord1, OGJK_REAL* d_coord2,        \
                                    void* d_simplices, OGJK_REAL* d_distances);                                \
  /* reference: allocate_epa_device_arrays / copy_epa_results_from_device / free_epa_device_arrays,            \
   * openGJK.h:155-194 (openGJK.cu:2956-2966, 3017-3032, 3050-3058) */                                         \
  int ogjk_##P##_allocate_epa_device_arrays(int n, OGJK_REAL** d_witness1, OGJK_REAL** d_witness2,             \
                                            OGJK_REAL** d_contact_normals);                                    \
  int ogjk_##P##_copy_epa_results_from_device(int n, const OGJK_REAL* d_witness1,                              \
                                              const OGJK_REAL* d_witness2,                                     \
                                              const OGJK_REAL* d_contact_normals, OGJK_REAL* witness1,         \
                                              OGJK_REAL* witness2, OGJK_REAL* contact_normals);                \
  int ogjk_##P##_free_epa_device_arrays(OGJK_REAL* d_witness1, OGJK_REAL* d_witness2,                          \
                                        OGJK_REAL* d_contact_normals);                                         \
  /* ---- indexed: one polytope pool + (idx1, idx2) pairs ------------------------------------------------ */ \
  /* reference: allocate_indexed_device, openGJK.h:330-340 (openGJK.cu:3193-3209); d_contact_normals may be    \
   * NULL; d_simplices is NOT zero-filled (as in the reference) */                                              \
  int ogjk_##P##_allocate_indexed_device(int num_polytopes, int max_pairs, const void* polytopes,              \
                                         void** d_polytopes, OGJK_REAL** d_coords, void** d_pairs,             \
                                         void** d_simplices, OGJK_REAL** d_distances,                          \
                                         OGJK_REAL** d_contact_normals);                                       \
  /* reference: free_indexed_device, openGJK.h:352-359 (openGJK.cu:3211-3225) */                                \
  int ogjk_##P##_free_indexed_device(void* d_polytopes, OGJK_REAL* d_coords, void* d_pairs,                    \
                                     void* d_simplices, OGJK_REAL* d_distances,                                \
                                     OGJK_REAL* d_contact_normals);                                            \
  /* reference: upload_pairs_device, openGJK.h:368-372 (openGJK.cu:3227-3233) */                                \
  int ogjk_##P##_upload_pairs_device(int num_pairs, const void* pairs, void* d_pairs);                         \
  /* reference: compute_minimum_distance_indexed, openGJK.h:398-405 (openGJK.cu:3064-3125) */                   \
  int ogjk_##P##_compute_minimum_distance_indexed(int num_polytopes, int num_pairs, const void* polytopes,     \
                                                  const void* pairs, void* simplices, OGJK_REAL* distances);   \
  /* reference: compute_minimum_distance_indexed_device, openGJK.h:419-425 (openGJK.cu:3127-3142) */            \
  int ogjk_##P##_compute_minimum_distance_indexed_device(int num_pairs, const void* d_polytopes,               \
                                                         const void* d_pairs, void* d_simplices,               \
                                                         OGJK_REAL* d_distances);                              \
  /* reference: compute_epa_indexed_device, openGJK.h:450-457 (openGJK.cu:3144-3163) */                         \
  int ogjk_##P##_compute_epa_indexed_device(int num_pairs, const void* d_polytopes, const void* d_pairs,       \
                                            void* d_simplices, OGJK_REAL* d_distances,                         \
                                            OGJK_REAL* d_contact_normals);                                     \
  /* the two calls above in one (the device part of compute_gjk_epa_indexed, openGJK.cu:3274-3311, i.e. what the  \
   * visualiser issues every physics step, integrate_final_gjk.cu:1028-1036); lets the library fuse the EPA gate    \
   * into the GJK kernel and feeds ogjk_stage_times. */                                                             \
  int ogjk_##P##_gjk_epa_indexed_device(int num_pairs, const void* d_polytopes, const void* d_pairs,           \
                                        void* d_simplices, OGJK_REAL* d_distances,                             \
                                        OGJK_REAL* d_contact_normals);                                         \
  /* reference: compute_epa_indexed, openGJK.h:473-481 (openGJK.cu:3235-3272) */                                \
  int ogjk_##P##_compute_epa_indexed(int num_polytopes, int num_pairs, const void* polytopes,                  \
                                     const void* pairs, void* simplices, OGJK_REAL* distances,                 \
                                     OGJK_REAL* contact_normals);                                              \
  /* reference: compute_gjk_epa_indexed, openGJK.h:497-505 (openGJK.cu:3274-3311) */                            \
  int ogjk_##P##_compute_gjk_epa_indexed(int num_polytopes, int num_pairs, const void* polytopes,              \
                                         const void* pairs, void* simplices, OGJK_REAL* distances,             \
                                         OGJK_REAL* contact_normals);                                          \
  /* ---- flat-array entry points (no descriptors): coords are n x nverts x 3, device resident ---------- */ \
  /* The same GJK / EPA, for callers (benchmarks, Python, other FFIs) that hold uniform batches as dense       \
   * arrays; equivalent to building descriptors with numpoints = nverts and coord = base + i*nverts*3. */       \
  int ogjk_##P##_gjk_uniform_device(int n, int nverts1, const OGJK_REAL* d_coord1, int nverts2,                \
                                    const OGJK_REAL* d_coord2, void* d_simplices, OGJK_REAL* d_distances);     \
  int ogjk_##P##_epa_uniform_device(int n, int nverts1, const OGJK_REAL* d_coord1, int nverts2,                \
                                    const OGJK_REAL* d_coord2, void* d_simplices, OGJK_REAL* d_distances,      \
                                    OGJK_REAL* d_contact_normals);                                             \
  /* Descriptor upkeep of a device-resident pool (reference visualization/integrate_final_gjk.cu:691-704             \
   * init_polytopes_kernel): polytopes[sm] = {numpoints = vert_counts[sm], coord = verts_world + 3*vert_offsets[sm],  \
   * s = 0, s_idx = 0}.  With NULL offsets/counts the pool is uniform (uniform_count vertices each, dense) and, when   \
   * uniform_count % 4 == 0, indexed batches over it take the slot kernels. */                                         \
  int ogjk_##P##_init_polytopes_device(void* d_polytopes, OGJK_REAL* d_verts_world, const int* d_vert_offsets,        \
                                       const int* d_vert_counts, int uniform_count, int num_submeshes);                \
  /* Contact response over the hot path's outputs (reference visualization/integrate_final_gjk.cu:572-689              \
   * collision_response_kernel and its call site :1039-1054): for every pair with distance <= params[0] the Baumgarte  \
   * position correction (distance < 0) and the normal impulse on linear / angular velocity, bodies = sub_mesh_body of  \
   * the pair's two indices (NULL: the indices themselves).  positions / vel / ang / quats: float4 per body (vel.w =    \
   * mass as in the reference), inv_inertia: 3 floats per body.  d_vel_pong / d_ang_pong receive ping + impulses for     \
   * EVERY body (the reference's ping -> pong cudaMemcpy is folded in); d_positions is corrected in place.              \
   * params = {epsilon, restitution, restitution_threshold, baumgarte_beta} (reference sim_config.h:60-64: 0.7, 2.0,    \
   * 0.2).  Deterministic: a body's contributions are added in ascending pair order, all pairs read the positions as    \
   * they were on entry (the reference's atomicAdd order is whatever the hardware makes it).  Physics state is fp32    \
   * in both precisions, as in the reference. */                                                                        \
  int ogjk_##P##_contact_response_device(int num_pairs, const void* d_pairs, const OGJK_REAL* d_distances,             \
                                         const void* d_simplices, const OGJK_REAL* d_contact_normals,                  \
                                         const int* d_sub_mesh_body, int num_objects, float* d_positions,              \
                                         const float* d_vel_ping, float* d_vel_pong, const float* d_ang_ping,          \
                                         float* d_ang_pong, const float* d_quats, const float* d_inv_inertia,          \
                                         const float* params);                                                         \
  /* GJK followed by EPA in one call (what compute_gjk_epa does after its upload, reference                   \
   * GJK/gpu/openGJK.cu:2854-2883); lets the library fuse the EPA gate into the GJK kernel. */                 \
  int ogjk_##P##_gjk_epa_uniform_device(int n, int nverts1, const OGJK_REAL* d_coord1, int nverts2,            \
                                        const OGJK_REAL* d_coord2, void* d_simplices, OGJK_REAL* d_distances,  \
                                        OGJK_REAL* d_contact_normals);

OGJK_DECLARE_API(f32, float)
OGJK_DECLARE_API(f64, double)

#ifdef __cplusplus
}
#endif
#endif /* OPENGJK_B200_H__ */

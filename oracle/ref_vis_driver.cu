/*
 * oracle/ref_vis_driver.cu -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Host drivers around the UNMODIFIED kernels of the reference's only real caller of the hot path,
 * /root/reference/visualization/integrate_final_gjk.cu: transform_to_world_kernel (:304-332), insert_objects /
 * count_pairs / generate_pairs (:467-570), collision_response_kernel (:572-689), init_polytopes_kernel (:691-704) and
 * the quat_rotate helpers (:102-118).  That file as a whole needs OpenGL headers and a window; the kernels do not.
 * oracle/build_ref_vis.sh therefore extracts exactly those function definitions -- by name, with awk, at build time,
 * into a scratch file that is deleted after compilation -- and compiles them together with this driver (which includes
 * the scratch file as REFVIS_KERNELS) with nvcc's default flags, as the reference's visualisation target does (only the
 * GJK library target carries --fmad=false, GJK/CMakeLists.txt:32).  Used to pin the numpy restatements of SURVEY.md
 * section 8(f) rows 1-3 (oracle/broadphase_oracle.py, transform_oracle.py, contact_oracle.py) and, on the GPU box, as a
 * live second opinion for the product's kernels.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>

#include <vector>

#include "GJK/gpu/openGJK.h"
#include "visualization/sim_config.h"

using std::max;
using std::min;

#include REFVIS_KERNELS

extern "C" {

/* sim_broad_phase (:916-1002) up to the body-pair list: memset, insert, count, exclusive scan (done on the host here:
 * integer prefix sums are exact), generate.  Returns the number of pairs found; at most max_pairs are written. */
long ogjk_refvis_broadphase(int n, const float* pos_radius, float cell_size, float boundary, int grid_size, int* pairs_out,
                            int max_pairs) {
  const int cells = grid_size * grid_size * grid_size;
  float4* d_pos;
  int *d_counts, *d_objs, *d_pc, *d_po;
  gkCollisionPair* d_pairs;
  cudaMalloc(&d_pos, n * sizeof(float4));
  cudaMalloc(&d_counts, cells * sizeof(int));
  cudaMalloc(&d_objs, (size_t)cells * MAX_OBJECTS_PER_CELL * sizeof(int));
  cudaMalloc(&d_pc, n * sizeof(int));
  cudaMalloc(&d_po, n * sizeof(int));
  cudaMalloc(&d_pairs, (size_t)(max_pairs > 0 ? max_pairs : 1) * sizeof(gkCollisionPair));
  cudaMemcpy(d_pos, pos_radius, n * sizeof(float4), cudaMemcpyHostToDevice);
  cudaMemset(d_counts, 0, cells * sizeof(int));
  const int blocks = (n + BLOCK_SIZE - 1) / BLOCK_SIZE;
  insert_objects_kernel<<<blocks, BLOCK_SIZE>>>(d_pos, n, d_counts, d_objs, cell_size, boundary, grid_size);
  cudaMemset(d_pc, 0, n * sizeof(int));
  count_pairs_kernel<<<blocks, BLOCK_SIZE>>>(n, d_counts, d_objs, d_pos, cell_size, boundary, d_pc, grid_size);
  std::vector<int> pc(n), po(n);
  cudaMemcpy(pc.data(), d_pc, n * sizeof(int), cudaMemcpyDeviceToHost);
  long total = 0;
  for (int i = 0; i < n; ++i) {
    po[i] = (int)total;
    total += pc[i];
  }
  cudaMemcpy(d_po, po.data(), n * sizeof(int), cudaMemcpyHostToDevice);
  if (total > 0 && max_pairs > 0) {
    generate_pairs_kernel<<<blocks, BLOCK_SIZE>>>(n, d_counts, d_objs, d_pos, cell_size, boundary, d_po, d_pairs, max_pairs,
                                                  grid_size);
    const long m = total < max_pairs ? total : max_pairs;
    cudaMemcpy(pairs_out, d_pairs, (size_t)m * sizeof(gkCollisionPair), cudaMemcpyDeviceToHost);
  }
  std::vector<int> cc(cells);
  cudaMemcpy(cc.data(), d_counts, cells * sizeof(int), cudaMemcpyDeviceToHost);
  int worst = 0;
  for (int c : cc) worst = c > worst ? c : worst;
  cudaFree(d_pos); cudaFree(d_counts); cudaFree(d_objs); cudaFree(d_pc); cudaFree(d_po); cudaFree(d_pairs);
  if (cudaGetLastError() != cudaSuccess) return -1;
  if (worst > MAX_OBJECTS_PER_CELL) return -2;  /* the reference silently dropped objects: not a usable comparison */
  return total;
}

int ogjk_refvis_transform(int num_submeshes, int num_bodies, int total_verts, const float* positions, const float* quats,
                          const float* scales, const float* verts_local, float* verts_world, const int* vert_offsets,
                          const int* vert_counts, const int* sub_mesh_body) {
  float4 *d_p, *d_q;
  float3 *d_s, *d_l, *d_w;
  int *d_o, *d_c, *d_b;
  cudaMalloc(&d_p, num_bodies * sizeof(float4));
  cudaMalloc(&d_q, num_bodies * sizeof(float4));
  cudaMalloc(&d_s, num_bodies * sizeof(float3));
  cudaMalloc(&d_l, (size_t)total_verts * sizeof(float3));
  cudaMalloc(&d_w, (size_t)total_verts * sizeof(float3));
  cudaMalloc(&d_o, num_submeshes * sizeof(int));
  cudaMalloc(&d_c, num_submeshes * sizeof(int));
  cudaMalloc(&d_b, num_submeshes * sizeof(int));
  cudaMemcpy(d_p, positions, num_bodies * sizeof(float4), cudaMemcpyHostToDevice);
  cudaMemcpy(d_q, quats, num_bodies * sizeof(float4), cudaMemcpyHostToDevice);
  cudaMemcpy(d_s, scales, num_bodies * sizeof(float3), cudaMemcpyHostToDevice);
  cudaMemcpy(d_l, verts_local, (size_t)total_verts * sizeof(float3), cudaMemcpyHostToDevice);
  cudaMemcpy(d_o, vert_offsets, num_submeshes * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(d_c, vert_counts, num_submeshes * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(d_b, sub_mesh_body, num_submeshes * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemset(d_w, 0, (size_t)total_verts * sizeof(float3));
  transform_to_world_kernel<<<(num_submeshes + BLOCK_SIZE - 1) / BLOCK_SIZE, BLOCK_SIZE>>>(d_p, d_q, d_s, d_l, d_w, d_o, d_c, d_b,
                                                                                         num_submeshes);
  cudaMemcpy(verts_world, d_w, (size_t)total_verts * sizeof(float3), cudaMemcpyDeviceToHost);
  cudaFree(d_p); cudaFree(d_q); cudaFree(d_s); cudaFree(d_l); cudaFree(d_w); cudaFree(d_o); cudaFree(d_c); cudaFree(d_b);
  return (int)cudaGetLastError();
}

/* collision_response_kernel + the ping -> pong copies of its call site (:1039-1054).  positions are corrected in place;
 * vel_pong / ang_pong receive ping + impulses. */
int ogjk_refvis_response(int num_pairs, const int* pairs, const float* distances, const void* simplices, const float* normals,
                         const int* sub_mesh_body, int num_submeshes, int num_objects, float* positions, const float* vel_ping,
                         float* vel_pong, const float* ang_ping, float* ang_pong, const float* quats,
                         const float* inv_inertia, float epsilon) {
  float4 *d_pos, *d_vp, *d_vq, *d_ap, *d_aq, *d_q;
  float3* d_ii;
  gkCollisionPair* d_pairs;
  gkFloat *d_dist, *d_nrm;
  gkSimplex* d_simp;
  int* d_smb;
  const size_t nb = (size_t)num_objects * sizeof(float4);
  cudaMalloc(&d_pos, nb); cudaMalloc(&d_vp, nb); cudaMalloc(&d_vq, nb); cudaMalloc(&d_ap, nb); cudaMalloc(&d_aq, nb);
  cudaMalloc(&d_q, nb);
  cudaMalloc(&d_ii, (size_t)num_objects * sizeof(float3));
  cudaMalloc(&d_pairs, (size_t)(num_pairs + 1) * sizeof(gkCollisionPair));
  cudaMalloc(&d_dist, (size_t)(num_pairs + 1) * sizeof(gkFloat));
  cudaMalloc(&d_nrm, (size_t)(num_pairs + 1) * 3 * sizeof(gkFloat));
  cudaMalloc(&d_simp, (size_t)(num_pairs + 1) * sizeof(gkSimplex));
  cudaMalloc(&d_smb, (size_t)num_submeshes * sizeof(int));
  cudaMemcpy(d_pos, positions, nb, cudaMemcpyHostToDevice);
  cudaMemcpy(d_vp, vel_ping, nb, cudaMemcpyHostToDevice);
  cudaMemcpy(d_ap, ang_ping, nb, cudaMemcpyHostToDevice);
  cudaMemcpy(d_q, quats, nb, cudaMemcpyHostToDevice);
  cudaMemcpy(d_ii, inv_inertia, (size_t)num_objects * sizeof(float3), cudaMemcpyHostToDevice);
  cudaMemcpy(d_pairs, pairs, (size_t)num_pairs * sizeof(gkCollisionPair), cudaMemcpyHostToDevice);
  cudaMemcpy(d_dist, distances, (size_t)num_pairs * sizeof(gkFloat), cudaMemcpyHostToDevice);
  cudaMemcpy(d_nrm, normals, (size_t)num_pairs * 3 * sizeof(gkFloat), cudaMemcpyHostToDevice);
  cudaMemcpy(d_simp, simplices, (size_t)num_pairs * sizeof(gkSimplex), cudaMemcpyHostToDevice);
  cudaMemcpy(d_smb, sub_mesh_body, (size_t)num_submeshes * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(d_vq, d_vp, nb, cudaMemcpyDeviceToDevice);
  cudaMemcpy(d_aq, d_ap, nb, cudaMemcpyDeviceToDevice);
  if (num_pairs > 0)
    collision_response_kernel<<<(num_pairs + BLOCK_SIZE - 1) / BLOCK_SIZE, BLOCK_SIZE>>>(
        d_pos, d_vp, d_vq, d_ap, d_aq, d_q, d_ii, d_pairs, d_dist, d_simp, d_nrm, d_smb, epsilon, num_pairs, num_objects);
  cudaMemcpy(positions, d_pos, nb, cudaMemcpyDeviceToHost);
  cudaMemcpy(vel_pong, d_vq, nb, cudaMemcpyDeviceToHost);
  cudaMemcpy(ang_pong, d_aq, nb, cudaMemcpyDeviceToHost);
  cudaFree(d_pos); cudaFree(d_vp); cudaFree(d_vq); cudaFree(d_ap); cudaFree(d_aq); cudaFree(d_q); cudaFree(d_ii);
  cudaFree(d_pairs); cudaFree(d_dist); cudaFree(d_nrm); cudaFree(d_simp); cudaFree(d_smb);
  return (int)cudaGetLastError();
}

/* init_polytopes_kernel: descriptors over a world-space pool; returned with coord expressed as an element offset */
int ogjk_refvis_init_polytopes(int num_submeshes, const int* vert_offsets, const int* vert_counts, int* numpoints_out,
                               long* coord_offset_out) {
  gkPolytope* d_poly;
  float3* d_w;
  int *d_o, *d_c;
  cudaMalloc(&d_poly, (size_t)num_submeshes * sizeof(gkPolytope));
  cudaMalloc(&d_w, 16);
  cudaMalloc(&d_o, num_submeshes * sizeof(int));
  cudaMalloc(&d_c, num_submeshes * sizeof(int));
  cudaMemcpy(d_o, vert_offsets, num_submeshes * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemcpy(d_c, vert_counts, num_submeshes * sizeof(int), cudaMemcpyHostToDevice);
  init_polytopes_kernel<<<(num_submeshes + BLOCK_SIZE - 1) / BLOCK_SIZE, BLOCK_SIZE>>>(d_poly, d_w, d_o, d_c, num_submeshes);
  std::vector<gkPolytope> h(num_submeshes);
  cudaMemcpy(h.data(), d_poly, (size_t)num_submeshes * sizeof(gkPolytope), cudaMemcpyDeviceToHost);
  for (int i = 0; i < num_submeshes; ++i) {
    numpoints_out[i] = h[i].numpoints;
    coord_offset_out[i] = (long)(h[i].coord - (gkFloat*)d_w);
  }
  cudaFree(d_poly); cudaFree(d_w); cudaFree(d_o); cudaFree(d_c);
  return (int)cudaGetLastError();
}

}  // extern "C"

// broadphase.cuh -- uniform-grid broad phase: bounding spheres -> gkCollisionPair list, on the device.
//
// the step immediately before the hot path in the reference's only real caller
// (visualization/integrate_final_gjk.cu: insert_objects_kernel :467-491, count_pairs_kernel :493-527,
// generate_pairs_kernel :529-570, driven by sim_broad_phase :916-1002).  Contract kept: objects are (x, y, z, bounding
// radius); an object lives in the cell floor((p + boundary) / cell_size), clamped to the grid; object i is paired with
// every object j > i found in the 27 cells around its own whose sphere overlaps, the test being
// fx*fx + fy*fy + fz*fz < (ri + rj)*(ri + rj) in fp32; pairs come out grouped by i in ascending order, each group at
// the exclusive prefix sum of the per-object counts, and writes beyond `max_pairs` are dropped.
// What differs is the machine mapping:
//   * cell lists are compact (histogram -> exclusive scan -> stable sort of the object ids by cell) instead of a fixed
//     512 ids per cell: 80 KB instead of 55 MB for the 30^3 grid, and no silent loss of the objects beyond 512 in a
//     crowded cell;
//   * counting and generation use one WARP per object: the lanes stride over the candidates of the 27 cells, counts
//     are reduced with REDUX, pairs are written through ballot/popc compaction (the reference walks ~1600 candidates
//     per thread serially with 20 000 threads in flight for BASELINE config 5);
//   * the arithmetic of the overlap test is written with explicitly rounded operations (the reference's build lets
//     nvcc contract it), so that the pair SET is reproducible and equals the numpy oracle bit for bit.
// The order of the pairs inside one object's group follows the cell lists.  The reference fills those with atomics, so
// its order changes from run to run; here every cell list is in ascending object id (the ids are sorted by cell with a
// stable radix sort -- cub::DeviceRadixSort, a library call off the hot path), so the whole pair list, and with it
// the pair-ordered contact response downstream, is reproducible run to run.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gjk_math.cuh"
#include "ogjk_types.h"

namespace ogjk {

OGJK_D int bp_cell_coord(float p, float boundary, float cell_size, int grid_size) {
  int c = (int)floorf(div_rn(add_rn(p, boundary), cell_size));
  c = c < 0 ? 0 : c;
  return c > grid_size - 1 ? grid_size - 1 : c;
}
OGJK_D int bp_cell_of(const float4& p, float boundary, float cell_size, int grid_size, int& cx, int& cy, int& cz) {
  cx = bp_cell_coord(p.x, boundary, cell_size, grid_size);
  cy = bp_cell_coord(p.y, boundary, cell_size, grid_size);
  cz = bp_cell_coord(p.z, boundary, cell_size, grid_size);
  return cx + cy * grid_size + cz * grid_size * grid_size;
}
OGJK_D bool bp_overlap(const float4& a, const float4& b) {
  const float fx = sub_rn(a.x, b.x), fy = sub_rn(a.y, b.y), fz = sub_rn(a.z, b.z);
  const float d2 = add_rn(add_rn(mul_rn(fx, fx), mul_rn(fy, fy)), mul_rn(fz, fz));
  const float rs = add_rn(a.w, b.w);
  return d2 < mul_rn(rs, rs);
}

// cell of every object + histogram
__global__ void __launch_bounds__(256)
bp_histogram_kernel(const float4* __restrict__ pos, int n, float cell_size, float boundary, int grid_size,
                    int* __restrict__ obj_cell, int* __restrict__ obj_id, int* __restrict__ cell_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cx, cy, cz;
  const int c = bp_cell_of(pos[i], boundary, cell_size, grid_size, cx, cy, cz);
  obj_cell[i] = c;
  obj_id[i] = i;
  atomicAdd(&cell_count[c], 1);
}

// Exclusive prefix sum of `in[0..n)` into `out[0..n]` (out[n] = total), one CTA of 1024 threads, any n: each round
// scans 1024 * kItems values (thread-local serial scan + shuffle/shared scan of the thread totals) and carries the sum.
__global__ void __launch_bounds__(1024)
bp_exclusive_scan_kernel(const int* __restrict__ in, int* __restrict__ out, int n) {
  constexpr int kItems = 8;
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024 * kItems) {
    int v[kItems];
    int sum = 0;
    const int first = base + tid * kItems;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      v[k] = first + k < n ? in[first + k] : 0;
      sum += v[k];
    }
    int incl = sum;  // inclusive scan of the thread totals
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_sums[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int carry = carry_s;
    int run = carry + (warp ? warp_sums[warp - 1] : 0) + incl - sum;  // exclusive prefix of this thread's first item
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      if (first + k < n) out[first + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_sums[31];
    __syncthreads();
  }
  if (tid == 0) out[n] = carry_s;
}

// One warp per object.  kWrite = false: pair_counts[obj] = number of partners; kWrite = true: the pairs themselves at
// pair_offsets[obj], dropped beyond max_pairs (reference generate_pairs_kernel :563-564).
template <bool kWrite>
__global__ void __launch_bounds__(256)
bp_pairs_kernel(const float4* __restrict__ pos, int n, float cell_size, float boundary, int grid_size,
                const int* __restrict__ cell_start, const int* __restrict__ cell_objs, int* __restrict__ pair_counts,
                const int* __restrict__ pair_offsets, CollisionPair* __restrict__ pairs, int max_pairs) {
  const int obj = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (obj >= n) return;
  const float4 p = pos[obj];
  int cx, cy, cz;
  bp_cell_of(p, boundary, cell_size, grid_size, cx, cy, cz);
  int count = 0;
  long long write = kWrite ? (long long)pair_offsets[obj] : 0;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int nx = cx + dx, ny = cy + dy, nz = cz + dz;
        if (nx < 0 || nx >= grid_size || ny < 0 || ny >= grid_size || nz < 0 || nz >= grid_size) continue;
        const int c = nx + ny * grid_size + nz * grid_size * grid_size;
        const int lo = cell_start[c], hi = cell_start[c + 1];
        for (int k0 = lo; k0 < hi; k0 += 32) {
          const int k = k0 + lane;
          int other = -1;
          bool hit = false;
          if (k < hi) {
            other = cell_objs[k];
            if (other > obj) hit = bp_overlap(p, pos[other]);
          }
          if (kWrite) {
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (hit) {
              const long long g = write + __popc(m & ((1u << lane) - 1u));
              if (g < max_pairs) {
                pairs[g].idx1 = obj;
                pairs[g].idx2 = other;
              }
            }
            write += __popc(m);
          } else {
            count += hit ? 1 : 0;
          }
        }
      }
  if (!kWrite) {
    count = __reduce_add_sync(0xffffffffu, count);
    if (lane == 0) pair_counts[obj] = count;
  }
}

}  // namespace ogjk

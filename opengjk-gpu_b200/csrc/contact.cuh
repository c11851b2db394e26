// contact.cuh -- contact response: the consumer of the hot path's outputs.
//
// the step right after GJK/EPA in the reference's caller
// (visualization/integrate_final_gjk.cu: collision_response_kernel :572-689, quat_rotate :102-118; constants
// visualization/sim_config.h:60-64).  Per colliding pair (distance <= epsilon) it reads the EPA normal, the two
// witness points and the (negative) distance, and applies
//   * a Baumgarte position correction to both bodies when distance < 0,
//   * a normal impulse (restitution above a closing-speed threshold) to the linear and angular velocities of both.
// The reference runs one thread per pair and scatters 18 float atomicAdd per pair into the body arrays, so the sum a
// body receives depends on the order the hardware happens to serialise the atomics in (and positions are read while
// other threads are still correcting them).  This version is deterministic:
//   1. cr_keys_kernel       one thread per pair: slot 2p (body A of pair p) and 2p+1 (body B) get the body id as sort
//                           key when the pair can contribute, a sentinel otherwise; per-body histogram
//   2. stable radix sort of the slots by body (cub::DeviceRadixSort, key bits = log2(bodies)): a body's slots end up
//                           contiguous and in ascending pair order
//   3. cr_accumulate_kernel one warp per body: 32 slots at a time, every lane evaluates the pair of its slot (all
//                           loads and the whole impulse computation run in parallel), then the warp folds the 32
//                           contributions into the body's state in slot order with shuffles
// which yields exactly what the reference's kernel gives when its atomics land in pair order and all position reads
// precede the corrections (one of the orders the reference itself may take).  Every fp32 operation is rounded
// separately in the source's association order (the CPU checker under tests/ restates it the same way).
// Contributions are recomputed in step 3 rather than stored in step 1: 48 bytes of GJK/EPA output are re-read per
// slot instead of writing and reading 2 x 36-byte records per pair.
#pragma once
#include <cuda_runtime.h>

#include "gjk_math.cuh"
#include "ogjk_types.h"
#include "transform.cuh"

namespace ogjk {

struct ContactParams {
  float epsilon;                // pairs with distance > epsilon are ignored (reference: params->collision_epsilon)
  float restitution;            // sim_config.h:60
  float restitution_threshold;  // sim_config.h:61
  float baumgarte_beta;         // sim_config.h:64
};

struct BodyDelta {  // what one pair adds to one body
  V3<float> dpos, dvel, dang;
  bool has_pos, has_vel;
};

OGJK_D V3<float> quat_rotate_inv_rn(const float4& q, const V3<float>& v) {
  return quat_rotate_rn(make_float4(-q.x, -q.y, -q.z, q.w), v);
}

// body ids of pair p, or false when the reference kernel returns before touching anything (:588-594)
template <typename T>
OGJK_D bool contact_candidate(int p, const CollisionPair* __restrict__ pairs, const T* __restrict__ distances,
                              const int* __restrict__ sub_mesh_body, float epsilon, int num_objects, int& idA,
                              int& idB) {
  if (distances[p] > (T)epsilon) return false;
  const int smA = pairs ? pairs[p].idx1 : p, smB = pairs ? pairs[p].idx2 : p;
  idA = sub_mesh_body ? sub_mesh_body[smA] : smA;
  idB = sub_mesh_body ? sub_mesh_body[smB] : smB;
  return !(idA < 0 || idA >= num_objects || idB < 0 || idB >= num_objects);
}

// The reference's per-pair arithmetic (:596-688) for the body on side `role` (0 = A, 1 = B).
template <typename T>
OGJK_D BodyDelta contact_delta(int p, int role, int idA, int idB, const float4* __restrict__ positions,
                               const float4* __restrict__ vel, const float4* __restrict__ ang,
                               const float4* __restrict__ quats, const float* __restrict__ inv_inertia,
                               const T* __restrict__ distances, const SimplexT<T>* __restrict__ simplices,
                               const T* __restrict__ normals, const ContactParams& prm) {
  BodyDelta out;
  out.has_pos = out.has_vel = false;
  out.dpos = out.dvel = out.dang = mk<float>(0.f, 0.f, 0.f);
  float nx = (float)normals[3 * (size_t)p], ny = (float)normals[3 * (size_t)p + 1], nz = (float)normals[3 * (size_t)p + 2];
  const float nlen = sqrt_rn(add_rn(add_rn(mul_rn(nx, nx), mul_rn(ny, ny)), mul_rn(nz, nz)));
  if (nlen < 0.0001f) return out;
  const float inv_n = div_rn(1.0f, nlen);
  nx = mul_rn(nx, inv_n);
  ny = mul_rn(ny, inv_n);
  nz = mul_rn(nz, inv_n);
  const V3<float> n = mk<float>(nx, ny, nz);
  const float4 posA = positions[idA], posB = positions[idB];
  const float4 velA = vel[idA], velB = vel[idB];
  const float inv_mA = div_rn(1.0f, velA.w), inv_mB = div_rn(1.0f, velB.w);
  const T dist = distances[p];
  if (dist < (T)0) {
    const float pen = (float)(-dist);
    const float corr = div_rn(mul_rn(prm.baumgarte_beta, pen), add_rn(inv_mA, inv_mB));
    const float k = role ? mul_rn(corr, inv_mB) : mul_rn(-corr, inv_mA);
    out.dpos = mk<float>(mul_rn(k, nx), mul_rn(k, ny), mul_rn(k, nz));
    out.has_pos = true;
  }
  const SimplexT<T>& s = simplices[p];
  const V3<float> rA = mk<float>(sub_rn((float)s.witnesses[0][0], posA.x), sub_rn((float)s.witnesses[0][1], posA.y),
                                 sub_rn((float)s.witnesses[0][2], posA.z));
  const V3<float> rB = mk<float>(sub_rn((float)s.witnesses[1][0], posB.x), sub_rn((float)s.witnesses[1][1], posB.y),
                                 sub_rn((float)s.witnesses[1][2], posB.z));
  const float4 aA = ang[idA], aB = ang[idB];
  const V3<float> vA_ang = cross(mk<float>(aA.x, aA.y, aA.z), rA);
  const V3<float> vB_ang = cross(mk<float>(aB.x, aB.y, aB.z), rB);
  const float rvx = sub_rn(add_rn(velB.x, vB_ang.x), add_rn(velA.x, vA_ang.x));
  const float rvy = sub_rn(add_rn(velB.y, vB_ang.y), add_rn(velA.y, vA_ang.y));
  const float rvz = sub_rn(add_rn(velB.z, vB_ang.z), add_rn(velA.z, vA_ang.z));
  const float vn = add_rn(add_rn(mul_rn(rvx, nx), mul_rn(rvy, ny)), mul_rn(rvz, nz));
  if (vn > 0.0f) return out;
  const float4 qA = quats[idA], qB = quats[idB];
  const V3<float> iA = mk<float>(inv_inertia[3 * idA], inv_inertia[3 * idA + 1], inv_inertia[3 * idA + 2]);
  const V3<float> iB = mk<float>(inv_inertia[3 * idB], inv_inertia[3 * idB + 1], inv_inertia[3 * idB + 2]);
  const V3<float> tA_body = quat_rotate_inv_rn(qA, cross(rA, n));
  const V3<float> tB_body = quat_rotate_inv_rn(qB, cross(rB, n));
  const V3<float> IA_tA = mk<float>(mul_rn(iA.x, tA_body.x), mul_rn(iA.y, tA_body.y), mul_rn(iA.z, tA_body.z));
  const V3<float> IB_tB = mk<float>(mul_rn(iB.x, tB_body.x), mul_rn(iB.y, tB_body.y), mul_rn(iB.z, tB_body.z));
  const float ang_denom_A = dot(tA_body, IA_tA), ang_denom_B = dot(tB_body, IB_tB);
  const float e = (-vn > prm.restitution_threshold) ? prm.restitution : 0.0f;
  const float j = div_rn(mul_rn(-add_rn(1.0f, e), vn),
                         add_rn(add_rn(add_rn(inv_mA, inv_mB), ang_denom_A), ang_denom_B));
  out.has_vel = true;
  if (role == 0) {
    const float mj = -j;
    out.dvel = mk<float>(mul_rn(mul_rn(mj, nx), inv_mA), mul_rn(mul_rn(mj, ny), inv_mA), mul_rn(mul_rn(mj, nz), inv_mA));
    out.dang = quat_rotate_rn(qA, mk<float>(mul_rn(mj, IA_tA.x), mul_rn(mj, IA_tA.y), mul_rn(mj, IA_tA.z)));
  } else {
    out.dvel = mk<float>(mul_rn(mul_rn(j, nx), inv_mB), mul_rn(mul_rn(j, ny), inv_mB), mul_rn(mul_rn(j, nz), inv_mB));
    out.dang = quat_rotate_rn(qB, mk<float>(mul_rn(j, IB_tB.x), mul_rn(j, IB_tB.y), mul_rn(j, IB_tB.z)));
  }
  return out;
}

// step 1: sort keys (body id, or num_objects for slots that cannot contribute), slot ids, per-body slot counts
template <typename T>
__global__ void __launch_bounds__(256)
cr_keys_kernel(const CollisionPair* __restrict__ pairs, const T* __restrict__ distances,
               const int* __restrict__ sub_mesh_body, float epsilon, int num_pairs, int num_objects,
               unsigned* __restrict__ keys, unsigned* __restrict__ slots, int* __restrict__ counts) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= num_pairs) return;
  int idA = 0, idB = 0;
  const bool ok = contact_candidate(p, pairs, distances, sub_mesh_body, epsilon, num_objects, idA, idB);
  uint2 k, s;
  k.x = ok ? (unsigned)idA : (unsigned)num_objects;
  k.y = ok ? (unsigned)idB : (unsigned)num_objects;
  s.x = 2u * (unsigned)p;
  s.y = 2u * (unsigned)p + 1u;
  reinterpret_cast<uint2*>(keys)[p] = k;
  reinterpret_cast<uint2*>(slots)[p] = s;
  if (ok) {
    atomicAdd(&counts[idA], 1);
    atomicAdd(&counts[idB], 1);
  }
}

// step 3: one warp per body.  positions_in/out may alias (a body's position is only written by its own warp, and
// read by others): to keep "all reads precede the corrections" the corrected positions go to positions_out, which the
// host swaps in afterwards when it had to be a separate buffer.
template <typename T>
__global__ void __launch_bounds__(256)
cr_accumulate_kernel(const float4* __restrict__ positions_in, float4* __restrict__ positions_out,
                     const float4* __restrict__ vel_ping, float4* __restrict__ vel_pong,
                     const float4* __restrict__ ang_ping, float4* __restrict__ ang_pong,
                     const float4* __restrict__ quats, const float* __restrict__ inv_inertia,
                     const CollisionPair* __restrict__ pairs, const T* __restrict__ distances,
                     const SimplexT<T>* __restrict__ simplices, const T* __restrict__ normals,
                     const int* __restrict__ sub_mesh_body, ContactParams prm, int num_objects,
                     const int* __restrict__ seg_start, const unsigned* __restrict__ sorted_slots) {
  const int body = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (body >= num_objects) return;
  const int begin = seg_start[body], end = seg_start[body + 1];
  float4 pos = positions_in[body], v = vel_ping[body], w = ang_ping[body];
  for (int base = begin; base < end; base += 32) {
    const int i = base + lane;
    BodyDelta d;
    d.has_pos = d.has_vel = false;
    d.dpos = d.dvel = d.dang = mk<float>(0.f, 0.f, 0.f);
    if (i < end) {
      const unsigned slot = sorted_slots[i];
      const int p = (int)(slot >> 1), role = (int)(slot & 1u);
      int idA = 0, idB = 0;
      contact_candidate(p, pairs, distances, sub_mesh_body, prm.epsilon, num_objects, idA, idB);
      d = contact_delta(p, role, idA, idB, positions_in, vel_ping, ang_ping, quats, inv_inertia, distances, simplices,
                        normals, prm);
    }
    const unsigned mp = __ballot_sync(0xffffffffu, d.has_pos), mv = __ballot_sync(0xffffffffu, d.has_vel);
    const int cnt = min(32, end - base);
    for (int k = 0; k < cnt; ++k) {  // fold in slot order; every lane keeps the same running sums
      if ((mp >> k) & 1u) {
        pos.x = add_rn(pos.x, __shfl_sync(0xffffffffu, d.dpos.x, k));
        pos.y = add_rn(pos.y, __shfl_sync(0xffffffffu, d.dpos.y, k));
        pos.z = add_rn(pos.z, __shfl_sync(0xffffffffu, d.dpos.z, k));
      }
      if ((mv >> k) & 1u) {
        v.x = add_rn(v.x, __shfl_sync(0xffffffffu, d.dvel.x, k));
        v.y = add_rn(v.y, __shfl_sync(0xffffffffu, d.dvel.y, k));
        v.z = add_rn(v.z, __shfl_sync(0xffffffffu, d.dvel.z, k));
        w.x = add_rn(w.x, __shfl_sync(0xffffffffu, d.dang.x, k));
        w.y = add_rn(w.y, __shfl_sync(0xffffffffu, d.dang.y, k));
        w.z = add_rn(w.z, __shfl_sync(0xffffffffu, d.dang.z, k));
      }
    }
  }
  if (lane == 0) {
    positions_out[body] = pos;
    vel_pong[body] = v;
    ang_pong[body] = w;
  }
}

}  // namespace ogjk

"""TEST INFRASTRUCTURE -- CPU restatement (numpy, fp32) of the reference's contact response.

Follows visualization/integrate_final_gjk.cu of the reference: collision_response_kernel :572-689 (per-pair Baumgarte
correction and normal impulse), quat_rotate / quat_rotate_inv :102-118, the ping -> pong copies at the call site
:1039-1054, constants visualization/sim_config.h:60-64.  Every fp32 operation is rounded separately in the source's
left-to-right order.  The reference scatters with float atomicAdd, i.e. in no defined order; this oracle fixes ONE of
the orders the reference may take: pairs in ascending index order (body A's terms before body B's), every pair reading
the positions as they were before any correction.  Only tests/ may import this module.
"""
import numpy as np

f32 = np.float32


def _quat_rotate(q, v):
    ux, uy, uz, s = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    vx, vy, vz = v[..., 0], v[..., 1], v[..., 2]
    dot_uv = (ux * vx + uy * vy) + uz * vz
    cx = uy * vz - uz * vy
    cy = uz * vx - ux * vz
    cz = ux * vy - uy * vx
    two_dot = f32(2.0) * dot_uv
    k = (f32(2.0) * s) * s - f32(1.0)
    two_s = f32(2.0) * s
    return np.stack([(two_dot * ux + k * vx) + two_s * cx,
                     (two_dot * uy + k * vy) + two_s * cy,
                     (two_dot * uz + k * vz) + two_s * cz], -1).astype(np.float32)


def _quat_rotate_inv(q, v):
    qc = q.copy()
    qc[..., :3] = -qc[..., :3]
    return _quat_rotate(qc, v)


def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], -1).astype(np.float32)


def _dot(a, b):
    return ((a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]).astype(np.float32)


def pair_terms(pairs, distances, witnesses, normals, sub_mesh_body, positions, vel, ang, quats, inv_inertia,
               epsilon=0.0, restitution=0.7, restitution_threshold=2.0, baumgarte_beta=0.2):
    """Per-pair contributions.  Returns dict with idA, idB [n] and, for side in A/B: dpos_*, dvel_*, dang_* [n,3] plus
    masks has_pos, has_vel [n] (a False mask means the reference thread returned before those atomics)."""
    pairs = np.asarray(pairs).reshape(-1, 2)  # gkCollisionPair = (idx1, idx2)
    n = len(pairs)
    num_objects = len(positions)
    sm_a, sm_b = pairs[:, 0].astype(np.int64), pairs[:, 1].astype(np.int64)
    dist = np.asarray(distances)
    live = ~(dist > dist.dtype.type(f32(epsilon)))
    if sub_mesh_body is None:
        id_a, id_b = sm_a.copy(), sm_b.copy()
    else:
        smb = np.asarray(sub_mesh_body)
        id_a = smb[np.where(live, sm_a, 0)].astype(np.int64)
        id_b = smb[np.where(live, sm_b, 0)].astype(np.int64)
    live &= ~((id_a < 0) | (id_a >= num_objects) | (id_b < 0) | (id_b >= num_objects))
    ia, ib = np.where(live, id_a, 0), np.where(live, id_b, 0)
    P, V, W, Q = (np.asarray(x, np.float32) for x in (positions, vel, ang, quats))
    I = np.asarray(inv_inertia, np.float32).reshape(-1, 3)
    with np.errstate(all="ignore"):
        nrm = np.asarray(normals).reshape(n, 3).astype(np.float32)
        nlen = np.sqrt((nrm[:, 0] * nrm[:, 0] + nrm[:, 1] * nrm[:, 1]) + nrm[:, 2] * nrm[:, 2]).astype(np.float32)
        live &= ~(nlen < f32(0.0001))
        inv_n = (f32(1.0) / nlen).astype(np.float32)
        nv = (nrm * inv_n[:, None]).astype(np.float32)
        wit = np.asarray(witnesses).reshape(n, 2, 3).astype(np.float32)
        r_a = (wit[:, 0] - P[ia, :3]).astype(np.float32)
        r_b = (wit[:, 1] - P[ib, :3]).astype(np.float32)
        inv_ma = (f32(1.0) / V[ia, 3]).astype(np.float32)
        inv_mb = (f32(1.0) / V[ib, 3]).astype(np.float32)
        has_pos = live & (dist < 0)
        pen = (-dist).astype(np.float32)
        corr = ((f32(baumgarte_beta) * pen) / (inv_ma + inv_mb)).astype(np.float32)
        dpos_a = (((-corr) * inv_ma)[:, None] * nv).astype(np.float32)
        dpos_b = ((corr * inv_mb)[:, None] * nv).astype(np.float32)
        va_ang = _cross(W[ia, :3], r_a)
        vb_ang = _cross(W[ib, :3], r_b)
        rv = ((V[ib, :3] + vb_ang) - (V[ia, :3] + va_ang)).astype(np.float32)
        vn = _dot(rv, nv)
        has_vel = live & ~(vn > 0)
        ta_body = _quat_rotate_inv(Q[ia], _cross(r_a, nv))
        tb_body = _quat_rotate_inv(Q[ib], _cross(r_b, nv))
        ia_ta = (I[ia] * ta_body).astype(np.float32)
        ib_tb = (I[ib] * tb_body).astype(np.float32)
        den_a, den_b = _dot(ta_body, ia_ta), _dot(tb_body, ib_tb)
        e = np.where(-vn > f32(restitution_threshold), f32(restitution), f32(0.0)).astype(np.float32)
        j = (((-(f32(1.0) + e)) * vn) / (((inv_ma + inv_mb) + den_a) + den_b)).astype(np.float32)
        mj = (-j).astype(np.float32)
        dvel_a = (((mj[:, None] * nv).astype(np.float32)) * inv_ma[:, None]).astype(np.float32)
        dvel_b = (((j[:, None] * nv).astype(np.float32)) * inv_mb[:, None]).astype(np.float32)
        dang_a = _quat_rotate(Q[ia], (mj[:, None] * ia_ta).astype(np.float32))
        dang_b = _quat_rotate(Q[ib], (j[:, None] * ib_tb).astype(np.float32))
    return dict(idA=ia, idB=ib, has_pos=has_pos, has_vel=has_vel, dpos_A=dpos_a, dpos_B=dpos_b, dvel_A=dvel_a,
                dvel_B=dvel_b, dang_A=dang_a, dang_B=dang_b)


def contact_response(pairs, distances, witnesses, normals, sub_mesh_body, positions, vel_ping, ang_ping, quats,
                     inv_inertia, **params):
    """-> (positions', vel_pong, ang_pong): float32 [bodies, 4] each, contributions added in ascending pair order"""
    t = pair_terms(pairs, distances, witnesses, normals, sub_mesh_body, positions, vel_ping, ang_ping, quats,
                   inv_inertia, **params)
    pos = np.array(positions, np.float32, copy=True)
    vel = np.array(vel_ping, np.float32, copy=True)
    ang = np.array(ang_ping, np.float32, copy=True)
    n = len(t["idA"])
    # interleave A and B terms: slot 2p = A of pair p, 2p+1 = B; ufunc.at applies the additions in index order
    ids = np.stack([t["idA"], t["idB"]], 1).reshape(-1)
    for arr, key, mask in ((pos, "dpos", t["has_pos"]), (vel, "dvel", t["has_vel"]), (ang, "dang", t["has_vel"])):
        d = np.stack([t[key + "_A"], t[key + "_B"]], 1).reshape(2 * n, 3)
        m = np.repeat(mask, 2)
        for c in range(3):
            np.add.at(arr[:, c], ids[m], d[m, c])
    return pos, vel, ang


def contact_response_loop(pairs, distances, witnesses, normals, sub_mesh_body, positions, vel_ping, ang_ping, quats,
                          inv_inertia, **params):
    """the same, with the accumulation as an explicit Python loop over pairs (small cases: checks the ufunc.at order)"""
    t = pair_terms(pairs, distances, witnesses, normals, sub_mesh_body, positions, vel_ping, ang_ping, quats,
                   inv_inertia, **params)
    pos = np.array(positions, np.float32, copy=True)
    vel = np.array(vel_ping, np.float32, copy=True)
    ang = np.array(ang_ping, np.float32, copy=True)
    for p in range(len(t["idA"])):
        for side, idx in (("A", t["idA"][p]), ("B", t["idB"][p])):
            if t["has_pos"][p]:
                pos[idx, :3] = pos[idx, :3] + t["dpos_" + side][p]
            if t["has_vel"][p]:
                vel[idx, :3] = vel[idx, :3] + t["dvel_" + side][p]
                ang[idx, :3] = ang[idx, :3] + t["dang_" + side][p]
    return pos, vel, ang

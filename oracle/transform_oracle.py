"""TEST INFRASTRUCTURE -- CPU restatement (numpy, fp32) of the reference's local -> world vertex transform.

Follows visualization/integrate_final_gjk.cu of the reference: transform_to_world_kernel :304-332 (scale, rotate,
translate) and quat_rotate :102-113, with every fp32 operation rounded separately in the source's left-to-right order.
Only tests/ may import this module.
"""
import numpy as np

f32 = np.float32


def quat_rotate(q, v):
    """q: [..., 4] (x, y, z, w), v: [..., 3]; float32 arrays"""
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


def transform_uniform(positions, quats, scales, verts_local):
    """positions [n,4], quats [n,4], scales [n,3], verts_local [n,V,3] -> world [n,V,3] (all float32)"""
    p = np.asarray(positions, np.float32)
    q = np.asarray(quats, np.float32)
    sc = np.asarray(scales, np.float32)
    lv = np.asarray(verts_local, np.float32) * sc[:, None, :]
    rv = quat_rotate(q[:, None, :], lv)
    return (rv + p[:, None, :3]).astype(np.float32)


def transform_ragged(positions, quats, scales, verts_local, offsets, counts, sub_body):
    out = np.zeros_like(np.asarray(verts_local, np.float32))
    for sm in range(len(offsets)):
        b = sub_body[sm]
        o, c = offsets[sm], counts[sm]
        out[o:o + c] = transform_uniform(positions[b:b + 1], quats[b:b + 1], scales[b:b + 1],
                                         np.asarray(verts_local, np.float32)[None, o:o + c])[0]
    return out

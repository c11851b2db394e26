"""TEST INFRASTRUCTURE -- CPU restatement (numpy, fp32) of the reference's uniform-grid broad phase.

Follows visualization/integrate_final_gjk.cu of the reference: cell of an object = floorf((p + boundary) / cell_size)
clamped to [0, grid_size) (insert_objects_kernel :478-485); object i is paired with every object j > i found in the 27
cells around its own cell (count_pairs_kernel / generate_pairs_kernel :510-524, :549-566) whose bounding sphere
overlaps: fx*fx + fy*fy + fz*fz < r_sum*r_sum in float, each operation rounded separately (the product's kernels use
explicitly rounded operations; the reference's build may contract them, which can only move pairs whose spheres
touch within one ulp).  The reference keeps at most 512 ids per cell (MAX_OBJECTS_PER_CELL, sim_config.h:16) and
silently drops the rest; neither this restatement nor the product does.  Only tests/ may import this module.
"""
import numpy as np


def cells(pos_radius, cell_size, boundary, grid_size):
    p = np.asarray(pos_radius, np.float32)
    f32 = np.float32
    c = np.floor((p[:, :3] + f32(boundary)) / f32(cell_size)).astype(np.int64)
    return np.clip(c, 0, grid_size - 1)


def pairs(pos_radius, cell_size, boundary, grid_size):
    """-> int32 [m, 2] sorted lexicographically (idx1 < idx2)"""
    p = np.ascontiguousarray(pos_radius, np.float32)
    n = p.shape[0]
    c = cells(p, cell_size, boundary, grid_size)
    key = (c[:, 2] * grid_size + c[:, 1]) * grid_size + c[:, 0]
    order = np.argsort(key, kind="stable")
    skey = key[order]
    out = []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                nc = c + np.array([dx, dy, dz])
                ok = np.all((nc >= 0) & (nc < grid_size), axis=1)
                nkey = (nc[:, 2] * grid_size + nc[:, 1]) * grid_size + nc[:, 0]
                lo = np.searchsorted(skey, nkey, "left")
                hi = np.searchsorted(skey, nkey, "right")
                cnt = np.where(ok, hi - lo, 0)
                src = np.repeat(np.arange(n), cnt)
                if src.size == 0:
                    continue
                start = np.repeat(lo, cnt)
                within = np.arange(src.size) - np.repeat(np.cumsum(cnt) - cnt, cnt)
                dst = order[start + within]
                m = dst > src
                src, dst = src[m], dst[m]
                a, b = p[src], p[dst]
                fx, fy, fz = a[:, 0] - b[:, 0], a[:, 1] - b[:, 1], a[:, 2] - b[:, 2]
                d2 = (fx * fx + fy * fy) + fz * fz          # float32 arrays: every operation rounds separately
                rs = a[:, 3] + b[:, 3]
                hit = d2 < rs * rs
                out.append(np.stack([src[hit], dst[hit]], 1))
    res = np.concatenate(out, 0) if out else np.zeros((0, 2), np.int64)
    res = res[np.lexsort((res[:, 1], res[:, 0]))]
    return res.astype(np.int32)


def brute_force(pos_radius):
    """all i < j with overlapping spheres (same fp32 test); equals pairs() when cell_size >= 2 * max radius and no
    object is clamped into a border cell from far outside the grid"""
    p = np.ascontiguousarray(pos_radius, np.float32)
    n = p.shape[0]
    i, j = np.triu_indices(n, 1)
    a, b = p[i], p[j]
    fx, fy, fz = a[:, 0] - b[:, 0], a[:, 1] - b[:, 1], a[:, 2] - b[:, 2]
    d2 = (fx * fx + fy * fy) + fz * fz
    rs = a[:, 3] + b[:, 3]
    hit = d2 < rs * rs
    return np.stack([i[hit], j[hit]], 1).astype(np.int32)

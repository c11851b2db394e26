"""Generates tests/golden/vis_reference_kernels.npz: outputs of the REFERENCE's own visualiser kernels (broad phase, world
transform, contact response, descriptor upkeep -- visualization/integrate_final_gjk.cu:304-332, 467-570, 572-704,
compiled unmodified by oracle/build_ref_vis.sh) on seeded inputs.  Needs a GPU and oracle/_ref_gpu/libogjk_refvis_f32.so:

    gpurun -- 'python tests/golden/make_vis_golden.py gpurun_out/vis_reference_kernels.npz'

The file pins the numpy restatements oracle/{broadphase,transform,contact}_oracle.py (tests/test_ref_vis_pinning.py,
CPU) so that SURVEY.md section 8(f) rows 1-3 are checked against the reference itself, not only against restatements.
Contact response: the reference adds with float atomics in whatever order the hardware takes, so its output is only
defined up to that order; the golden holds one run.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from _pkgpath import load_oracle  # noqa: E402

SIMPLEX = np.dtype({"names": ["nvrtx", "vrtx", "vrtx_idx", "witnesses"],
                    "formats": ["<i4", ("<f4", (4, 3)), ("<i4", (4, 2)), ("<f4", (2, 3))],
                    "offsets": [0, 4, 52, 84], "itemsize": 108})


def scene(n, seed, boundary=12.0, rmin=0.3, rmax=1.4):
    rng = np.random.default_rng(seed)
    p = np.empty((n, 4), np.float32)
    p[:, :3] = rng.uniform(-boundary, boundary, size=(n, 3))
    p[:, 3] = rng.uniform(rmin, rmax, size=n)
    return p


def bodies(n, seed):
    rng = np.random.default_rng(seed)
    pos = np.zeros((n, 4), np.float32)
    pos[:, :3] = rng.uniform(-20, 20, (n, 3))
    pos[:, 3] = rng.uniform(0.3, 2.5, n)
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    sc = rng.uniform(0.2, 3.0, (n, 3)).astype(np.float32)
    return pos, q.astype(np.float32), sc


def contact_state(nb, seed, box=4.0):
    rng = np.random.default_rng(seed)
    pos = np.zeros((nb, 4), np.float32)
    pos[:, :3] = rng.uniform(-box, box, (nb, 3))
    pos[:, 3] = rng.uniform(0.3, 1.5, nb)
    vel = np.zeros((nb, 4), np.float32)
    vel[:, :3] = rng.normal(0, 3.0, (nb, 3))
    vel[:, 3] = rng.uniform(0.5, 4.0, nb)
    ang = np.zeros((nb, 4), np.float32)
    ang[:, :3] = rng.normal(0, 1.0, (nb, 3))
    q = rng.standard_normal((nb, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    inv_i = rng.uniform(0.2, 3.0, (nb, 3)).astype(np.float32)
    return pos, vel, ang, q.astype(np.float32), inv_i


def contacts(npairs, nsub, seed):
    rng = np.random.default_rng(seed)
    pairs = rng.integers(0, nsub, (npairs, 2)).astype(np.int32)
    same = pairs[:, 0] == pairs[:, 1]
    pairs[same, 1] = (pairs[same, 0] + 1) % nsub
    kind = rng.integers(0, 4, npairs)
    dist = np.where(kind == 0, rng.uniform(0.01, 2.0, npairs), np.where(kind == 1, 0.0, -rng.uniform(1e-4, 0.4, npairs)))
    nrm = rng.standard_normal((npairs, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm[rng.random(npairs) < 0.02] = 0.0
    wit = rng.uniform(-4.0, 4.0, (npairs, 2, 3))
    return pairs, dist.astype(np.float32), wit.astype(np.float32), nrm.astype(np.float32)


def main(path):
    rv = load_oracle().RefVis()
    out = {}
    # ---- broad phase: three scenes (cells >= sphere diameter; cells smaller than the spheres + clamped objects; tiny)
    for tag, (n, seed, bnd_scene, cell, boundary, grid) in {"bp0": (1500, 1, 12.0, 2.8, 12.0, 9), "bp1": (800, 2, 15.0, 1.0, 12.0, 24),
                                                           "bp2": (300, 3, 13.2, 5.0, 12.0, 5)}.items():
        p = scene(n, seed, boundary=bnd_scene)
        pairs, total = rv.broadphase(p, cell, boundary, grid, 400000)
        assert total == pairs.shape[0]
        pairs = pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]
        out[f"{tag}_pos"] = p
        out[f"{tag}_prm"] = np.array([cell, boundary, grid], np.float64)
        out[f"{tag}_pairs"] = pairs
        print(tag, n, "->", total, "pairs")
    # ---- world transform: ragged sub-meshes, several per body
    nb = 400
    pos, q, sc = bodies(nb, 3)
    rng = np.random.default_rng(4)
    counts = rng.integers(4, 40, 900).astype(np.int32)
    offsets = (np.cumsum(counts) - counts).astype(np.int32)
    sub_body = rng.integers(0, nb, 900).astype(np.int32)
    local = rng.standard_normal((int(counts.sum()), 3)).astype(np.float32)
    world = rv.transform(pos, q, sc, local, offsets, counts, sub_body)
    out.update(tr_pos=pos, tr_quat=q, tr_scale=sc, tr_local=local, tr_offsets=offsets, tr_counts=counts, tr_sub_body=sub_body,
               tr_world=world)
    npts, coff = rv.init_polytopes(offsets, counts)
    out.update(ip_numpoints=npts, ip_coord_offset=coff)
    # ---- contact response.  The reference reads positions while other threads atomicAdd corrections to them and adds
    # impulses with float atomics, so with several contacts per body its output depends on thread timing.  The pinned
    # scene therefore has DISJOINT pairs (every body in exactly one pair, through a sub-mesh -> body map with a few
    # invalid owners): there the reference's output is a pure function of its input.
    nb = 600
    npairs = nb // 2
    rng = np.random.default_rng(3)
    perm = rng.permutation(nb).astype(np.int32)
    smb = perm.copy()                       # sub-mesh s belongs to body perm[s]
    smb[rng.random(nb) < 0.03] = -1
    smb[rng.random(nb) < 0.03] = nb + 4
    cpos, vel, ang, cq, inv_i = contact_state(nb, 8)
    pairs, dist, wit, nrm = contacts(npairs, nb, 9)
    pairs = np.stack([np.arange(0, nb, 2), np.arange(1, nb, 2)], 1).astype(np.int32)   # sub-mesh pairs (2k, 2k+1)
    simp = np.zeros(npairs, SIMPLEX)
    simp["witnesses"] = wit
    eps = 0.1  # COLLISION_EPSILON, sim_config.h
    p2, v2, a2 = rv.response(pairs, dist, simp, nrm, smb, cpos, vel, ang, cq, inv_i, eps)
    out.update(cr_pairs=pairs, cr_dist=dist, cr_wit=wit, cr_nrm=nrm, cr_sub_body=smb, cr_pos=cpos, cr_vel=vel, cr_ang=ang,
               cr_quat=cq, cr_inv_inertia=inv_i, cr_eps=np.float32(eps), cr_pos_out=p2, cr_vel_out=v2, cr_ang_out=a2)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "vis_reference_kernels.npz"))

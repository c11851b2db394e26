"""Pins the checkers of SURVEY.md section 8(f) rows 1-3 to the REFERENCE ITSELF.

The numpy restatements (oracle/broadphase_oracle.py, transform_oracle.py, contact_oracle.py) used to be checked only
against brute force / their own loops.  Here they -- and on the GPU box the product's kernels -- are compared with
outputs of the reference visualiser's own kernels (visualization/integrate_final_gjk.cu:304-332, 467-570, 572-704,
compiled unmodified by oracle/build_ref_vis.sh):
  * CPU tests read tests/golden/vis_reference_kernels.npz, generated on a B200 by tests/golden/make_vis_golden.py;
  * GPU tests run the reference kernels live (oracle/_ref_gpu/libogjk_refvis_f32.so travels to the box).
What "equal" means per kernel: broad phase -- the same pair SET (the reference's order inside an object's group
follows atomic slot order) except pairs whose spheres touch within rounding, because the reference's visualiser is
built with FMA contraction on; world transform -- within 4 ulp-ish relative 1e-6 of the reference (same contraction
caveat; the fraction of bit-identical coordinates is reported); contact response -- within 1e-5 on scenes whose pairs
are disjoint (with several contacts per body the reference reads positions that other threads are correcting and adds
with float atomics, so its own output is timing-dependent); descriptor upkeep -- identical.
"""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT

GOLDEN = os.path.join(ROOT, "tests", "golden", "vis_reference_kernels.npz")


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "oracle", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _golden():
    if not os.path.exists(GOLDEN):
        pytest.skip("tests/golden/vis_reference_kernels.npz not generated yet (needs a GPU: tests/golden/make_vis_golden.py)")
    return np.load(GOLDEN)


def _pair_sets_agree(p, got, want):
    """identical, or differing only in pairs whose |c|^2 is within 4 ulp of r^2 (FMA contraction in the reference)"""
    a = {tuple(x) for x in np.asarray(got).tolist()}
    b = {tuple(x) for x in np.asarray(want).tolist()}
    diff = a ^ b
    for i, j in diff:
        d = p[i, :3].astype(np.float64) - p[j, :3].astype(np.float64)
        d2, r2 = float(d @ d), float(p[i, 3] + p[j, 3]) ** 2
        assert abs(d2 - r2) <= 4 * np.finfo(np.float32).eps * r2, (i, j, d2, r2)
    return len(diff)


@pytest.mark.parametrize("tag", ["bp0", "bp1", "bp2"])
def test_broadphase_oracle_against_reference_golden(tag):
    g = _golden()
    bp = _load("broadphase_oracle")
    p = g[f"{tag}_pos"]
    cell, boundary, grid = g[f"{tag}_prm"]
    got = bp.pairs(p, float(cell), float(boundary), int(grid))
    want = g[f"{tag}_pairs"]
    assert want.shape[0] > 50
    ndiff = _pair_sets_agree(p, got, want)
    assert ndiff <= max(2, want.shape[0] // 2000)


def test_transform_oracle_against_reference_golden():
    g = _golden()
    tr = _load("transform_oracle")
    got = tr.transform_ragged(g["tr_pos"], g["tr_quat"], g["tr_scale"], g["tr_local"], g["tr_offsets"], g["tr_counts"],
                              g["tr_sub_body"])
    want = g["tr_world"]
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6)
    off = (np.cumsum(g["tr_counts"]) - g["tr_counts"]).astype(np.int64)
    assert np.array_equal(g["ip_numpoints"], g["tr_counts"]) and np.array_equal(g["ip_coord_offset"], 3 * off)


def test_contact_oracle_against_reference_golden():
    g = _golden()
    co = _load("contact_oracle")
    got = co.contact_response(g["cr_pairs"], g["cr_dist"], g["cr_wit"], g["cr_nrm"], g["cr_sub_body"], g["cr_pos"], g["cr_vel"],
                              g["cr_ang"], g["cr_quat"], g["cr_inv_inertia"], epsilon=float(g["cr_eps"]))
    for name, a, b in zip(("positions", "velocities", "angular"), got, (g["cr_pos_out"], g["cr_vel_out"], g["cr_ang_out"])):
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-5, err_msg=name)
    assert not np.array_equal(g["cr_vel_out"], g["cr_vel"])  # the scene does produce impulses
    assert not np.array_equal(g["cr_pos_out"], g["cr_pos"])  # ... and Baumgarte corrections


# ---- live, on the GPU box: the product's kernels against the reference's kernels --------------------------------------
def _refvis(oracle_mod):
    if not oracle_mod.RefVis.available():
        pytest.skip("oracle/_ref_gpu/libogjk_refvis_f32.so not built")
    return oracle_mod.RefVis()


@pytest.mark.gpu
@pytest.mark.parametrize("n,cell,boundary,grid", [(20000, 2.8, 12.0, 9), (5000, 1.0, 12.0, 24), (300, 5.0, 12.0, 5)])
def test_device_broadphase_against_reference_kernels(pkg, oracle_mod, n, cell, boundary, grid):
    import torch
    rv = _refvis(oracle_mod)
    rng = np.random.default_rng(3)
    p = np.empty((n, 4), np.float32)
    p[:, :3] = rng.uniform(-boundary * 1.1, boundary * 1.1, size=(n, 3))
    p[:, 3] = rng.uniform(0.3, 1.4, size=n)
    want, total = rv.broadphase(p, cell, boundary, grid, 4_000_000)
    assert total == want.shape[0]
    eng = pkg.Engine(np.float32)
    cap = total + 64
    d_pairs = torch.full((cap, 2), -1, dtype=torch.int32, device="cuda")
    mine = eng.broadphase_pairs_device(n, torch.from_numpy(p).cuda(), cell, boundary, grid, d_pairs, cap)
    torch.cuda.synchronize()
    got = d_pairs.cpu().numpy()[:mine]
    ndiff = _pair_sets_agree(p, got, want)
    assert ndiff <= max(2, total // 2000)
    # same grouping contract as the reference: ascending idx1, idx1 < idx2
    assert np.all(np.diff(want[:, 0]) >= 0) and np.all(np.diff(got[:, 0]) >= 0)


@pytest.mark.gpu
def test_device_transform_against_reference_kernels(pkg, oracle_mod):
    import torch
    rv = _refvis(oracle_mod)
    rng = np.random.default_rng(4)
    nb, nsub = 3000, 5000
    pos = np.zeros((nb, 4), np.float32)
    pos[:, :3] = rng.uniform(-20, 20, (nb, 3))
    q = rng.standard_normal((nb, 4))
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    sc = rng.uniform(0.2, 3.0, (nb, 3)).astype(np.float32)
    counts = rng.integers(4, 40, nsub).astype(np.int32)
    offsets = (np.cumsum(counts) - counts).astype(np.int32)
    sub_body = rng.integers(0, nb, nsub).astype(np.int32)
    local = rng.standard_normal((int(counts.sum()), 3)).astype(np.float32)
    want = rv.transform(pos, q, sc, local, offsets, counts, sub_body)
    eng = pkg.Engine(np.float32)
    d_out = torch.zeros(local.shape, dtype=torch.float32, device="cuda")
    eng.transform_to_world_device(nsub, torch.from_numpy(pos).cuda(), torch.from_numpy(q).cuda(), torch.from_numpy(sc).cuda(),
                                  torch.from_numpy(local).cuda(), d_out, torch.from_numpy(offsets).cuda(),
                                  torch.from_numpy(counts).cuda(), torch.from_numpy(sub_body).cuda())
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6)
    print("world transform: bit-identical to the reference kernel on", float(np.mean(got == want)), "of the coordinates")


@pytest.mark.gpu
def test_device_contact_response_against_reference_kernels(pkg, oracle_mod):
    import torch
    from test_contact import _contacts, _run_device, _state
    rv = _refvis(oracle_mod)
    nb = 20000  # disjoint pairs (see the module docstring): sub-mesh pairs (2k, 2k+1), sub-mesh s -> body perm[s]
    npairs = nb // 2
    rng = np.random.default_rng(3)
    smb = rng.permutation(nb).astype(np.int32)
    smb[rng.random(nb) < 0.03] = -1
    pos, vel, ang, q, inv_i = _state(nb, 8)
    _p, dist, wit, nrm = _contacts(npairs, nb, 9)
    pairs = np.stack([np.arange(0, nb, 2), np.arange(1, nb, 2)], 1).astype(np.int32)
    simp = np.zeros(npairs, pkg.simplex_dtype(np.float32))
    simp["witnesses"] = wit
    want = rv.response(pairs, dist, simp, nrm, smb, pos, vel, ang, q, inv_i, 0.1)
    got = _run_device(pkg, np.float32, pairs, dist, wit, nrm, smb, pos, vel, ang, q, inv_i, epsilon=0.1)
    for name, a, b in zip(("positions", "velocities", "angular"), got, want):
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-5, err_msg=name)
    print("contact response: bit-identical to the reference kernel on",
          [float(np.mean(a == b)) for a, b in zip(got, want)], "of positions / velocities / angular velocities")

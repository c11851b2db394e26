"""Broad phase: the numpy oracle against brute force on the CPU; the CUDA kernels against
the oracle, and the whole device pipeline broad phase -> indexed GJK -> indexed EPA against the CPU oracle, on the GPU."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT, live_simplex_equal


def _bp_oracle():
    spec = importlib.util.spec_from_file_location("broadphase_oracle", os.path.join(ROOT, "oracle", "broadphase_oracle.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _scene(n, seed, boundary=12.0, rmin=0.3, rmax=1.4):
    rng = np.random.default_rng(seed)
    p = np.empty((n, 4), np.float32)
    p[:, :3] = rng.uniform(-boundary, boundary, size=(n, 3))
    p[:, 3] = rng.uniform(rmin, rmax, size=n)
    return p


def test_oracle_equals_brute_force():
    bp = _bp_oracle()
    p = _scene(1500, 1)
    cell = 2.8  # >= 2 * max radius: the 27-cell neighbourhood sees every overlapping pair
    grid = int(np.ceil(24.0 / cell))
    got = bp.pairs(p, cell, 12.0, grid)
    want = bp.brute_force(p)
    assert got.shape[0] > 500
    assert np.array_equal(got, want)


def test_oracle_clamps_and_small_cells():
    bp = _bp_oracle()
    p = _scene(800, 2, boundary=15.0)          # some objects lie outside the +-12 grid: clamped into border cells
    got = bp.pairs(p, 1.0, 12.0, 24)           # cells smaller than the spheres: a subset of the overlapping pairs
    allp = bp.brute_force(p)
    s_all = {tuple(x) for x in allp.tolist()}
    assert 0 < got.shape[0] < allp.shape[0]
    assert all(tuple(x) in s_all for x in got.tolist())
    c = bp.cells(p, 1.0, 12.0, 24)
    assert c.min() == 0 and c.max() == 23


@pytest.mark.gpu
@pytest.mark.parametrize("n,cell,boundary,grid", [(20000, 2.8, 12.0, 9), (5000, 1.0, 12.0, 24), (300, 5.0, 12.0, 5),
                                                  (40000, 3.0, 30.0, 20)])
def test_device_broadphase_matches_oracle(pkg, n, cell, boundary, grid):
    import torch
    bp = _bp_oracle()
    p = _scene(n, 3, boundary=boundary * 1.1)
    want = bp.pairs(p, cell, boundary, grid)
    eng = pkg.Engine(np.float32)
    d_p = torch.from_numpy(p).cuda()
    cap = max(1, want.shape[0] + 7)
    d_pairs = torch.full((cap, 2), -1, dtype=torch.int32, device="cuda")
    total = eng.broadphase_pairs_device(n, d_p, cell, boundary, grid, d_pairs, cap)
    torch.cuda.synchronize()
    assert total == want.shape[0]
    got = d_pairs.cpu().numpy()[:total]
    assert np.all(got[:, 0] < got[:, 1])
    assert np.all(np.diff(got[:, 0]) >= 0), "pairs must be grouped by idx1 in ascending order"
    raw = got.copy()
    got = got[np.lexsort((got[:, 1], got[:, 0]))]
    assert np.array_equal(got, want)
    assert np.all(d_pairs.cpu().numpy()[total:] == -1)
    # reproducible ORDER too (cell lists are sorted by object id, not filled through an atomic cursor): the contact
    # response downstream folds contributions in pair order
    for _ in range(2):
        d_again = torch.full((cap, 2), -1, dtype=torch.int32, device="cuda")
        assert eng.broadphase_pairs_device(n, d_p, cell, boundary, grid, d_again, cap) == total
        torch.cuda.synchronize()
        assert np.array_equal(d_again.cpu().numpy()[:total], raw)
    # clamp: a buffer that is too small is filled and nothing is written past it
    small = total // 2
    d_small = torch.full((small + 5, 2), -1, dtype=torch.int32, device="cuda")
    total2 = eng.broadphase_pairs_device(n, d_p, cell, boundary, grid, d_small, small)
    torch.cuda.synchronize()
    assert total2 == total
    tail = d_small.cpu().numpy()
    assert np.all(tail[small:] == -1) and np.all(tail[:small, 0] >= 0)


@pytest.mark.gpu
def test_device_pipeline_broadphase_gjk_epa(pkg, oracle_mod):
    """BASELINE config 5 end to end on the device: bounding spheres -> candidate pairs -> indexed GJK -> indexed EPA"""
    import torch
    npoly, nv = 5000, 32
    rng = np.random.default_rng(11)
    local = pkg.workloads.unit_sphere_hulls(npoly, nv, 5, np.float64)
    radius = rng.uniform(0.3, 1.4, npoly)
    centre = rng.uniform(-9.0, 9.0, (npoly, 3))
    pool = (local * radius[:, None, None] + centre[:, None, :]).astype(np.float32)
    spheres = np.concatenate([centre, radius[:, None] * 1.0001], 1).astype(np.float32)
    bp = _bp_oracle()
    cell, boundary, grid = 2.9, 10.0, 7
    want_pairs = bp.pairs(spheres, cell, boundary, grid)
    assert want_pairs.shape[0] >= 32768  # enough for the slot kernels to take the indexed batch
    eng = pkg.Engine(np.float32)
    desc, _keep = pkg.make_polytopes(pool)
    cap = want_pairs.shape[0]
    dp, dc, dpairs, dsimp, ddist, dnrm = eng.allocate_indexed_device(desc, cap)
    try:
        total = eng.broadphase_pairs_device(npoly, torch.from_numpy(spheres).cuda(), cell, boundary, grid, dpairs, cap)
        assert total == cap
        eng.compute_minimum_distance_indexed_device(total, dp, dpairs, dsimp, ddist)
        eng.compute_epa_indexed_device(total, dp, dpairs, dsimp, ddist, dnrm)
        simp, dist = eng.copy_results_from_device(total, dsimp, ddist)
        got_pairs = torch.empty((total, 2), dtype=torch.int32, device="cuda")
        import ctypes
        ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(got_pairs.data_ptr()), ctypes.c_void_p(dpairs),
                                               ctypes.c_size_t(8 * total), ctypes.c_int(3))
        got_pairs = got_pairs.cpu().numpy()
    finally:
        eng.free_indexed_device(dp, dc, dpairs, dsimp, ddist, dnrm)
    off = np.arange(npoly + 1) * nv
    es, ed, _en = oracle_mod.Oracle("port", np.float32).gjk_epa_indexed(pool.reshape(-1, 3), got_pairs, off, nthreads=8)
    assert np.array_equal(dist, ed) and live_simplex_equal(simp, es)
    order = np.lexsort((got_pairs[:, 1], got_pairs[:, 0]))
    assert np.array_equal(got_pairs[order], want_pairs)

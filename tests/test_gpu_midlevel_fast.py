"""The reference's main device-level sequence -- allocate_and_copy_device_arrays + compute_minimum_distance_device
(+ compute_epa_device) + copy_results_from_device, i.e. what GJK::GPU::computeDistances does (reference
examples/gpu/example.cu:23-52) -- on batches large enough for the slot kernels: arrays uploaded by the library are
remembered, so the descriptor-based call takes the dense fast kernels.  Bit-exact against the oracle."""
import numpy as np
import pytest

from conftest import live_simplex_equal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("nverts,spread", [(64, 10.0), (32, 1.5), (50, 6.0), (96, 8.0)])
def test_mid_level_sequence_large_batch(pkg, oracle_mod, dtype, nverts, spread):
    n = 40000
    a, b = pkg.workloads.random_pairs(n, nverts, spread, seed=606, dtype=dtype)
    eng = pkg.Engine(dtype)
    bd1, _k1 = pkg.make_polytopes(a)
    bd2, _k2 = pkg.make_polytopes(b)
    h = eng.allocate_and_copy_device_arrays(bd1, bd2)
    d_bd1, d_bd2, d_c1, d_c2, d_simp, d_dist = h
    d_w1, d_w2, d_nrm = eng.allocate_epa_device_arrays(n)
    try:
        eng.compute_minimum_distance_device(n, d_bd1, d_bd2, d_simp, d_dist)
        simp, dist = eng.copy_results_from_device(n, d_simp, d_dist)
        orc = oracle_mod.Oracle("port", dtype)
        os_, od = orc.gjk(a, b, nthreads=8)
        assert np.array_equal(dist, od) and live_simplex_equal(simp, os_)
        eng.compute_epa_device(n, d_bd1, d_bd2, d_simp, d_dist, d_nrm)
        simp2, dist2 = eng.copy_results_from_device(n, d_simp, d_dist)
        es, ed, _en = orc.epa(a, b, os_, od, nthreads=8)
        assert np.array_equal(dist2, ed) and live_simplex_equal(simp2, es)
    finally:
        eng.free_epa_device_arrays(d_w1, d_w2, d_nrm)
        eng.free_device_arrays(*h)


def test_ragged_batch_still_general_path(pkg, oracle_mod):
    rng = np.random.default_rng(8)
    counts = rng.integers(3, 70, size=3000)
    a = [pkg.workloads.random_polytopes(1, int(c), 6.0, 10 + i, np.float32)[0] for i, c in enumerate(counts)]
    b = [pkg.workloads.random_polytopes(1, int(c), 6.0, 90000 + i, np.float32)[0] for i, c in enumerate(counts[::-1])]
    eng = pkg.Engine(np.float32)
    bd1, _k1 = pkg.make_polytopes(a)
    bd2, _k2 = pkg.make_polytopes(b)
    h = eng.allocate_and_copy_device_arrays(bd1, bd2)
    try:
        eng.compute_minimum_distance_device(len(a), h[0], h[1], h[4], h[5])
        simp, dist = eng.copy_results_from_device(len(a), h[4], h[5])
    finally:
        eng.free_device_arrays(*h)
    off1 = np.concatenate([[0], np.cumsum([len(x) for x in a])])
    off2 = np.concatenate([[0], np.cumsum([len(x) for x in b])])
    os_, od = oracle_mod.Oracle("port", np.float32).gjk(np.concatenate(a), np.concatenate(b), off1, off2)
    assert np.array_equal(dist, od) and live_simplex_equal(simp, os_)

"""GPU parity through the explicit-device-memory and indexed entry points (reference GJK/gpu/openGJK.h:155-505,
the API self-consistency tests 1-6 of examples/main.cpp:378-644) and the flat uniform fast path."""
import numpy as np
import pytest

from conftest import live_simplex_equal

pytestmark = pytest.mark.gpu


def _oracle(oracle_mod, dtype, a, b):
    orc = oracle_mod.Oracle("port", dtype)
    s, d = orc.gjk(a, b)
    es, ed, en = orc.epa(a, b, s, d)
    return (s, d), (es, ed, en)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_mid_level_device_api(pkg, oracle_mod, dtype):
    import torch
    a, b = pkg.workloads.random_pairs(12000, 24, 1.5, seed=11, dtype=dtype)
    eng = pkg.Engine(dtype)
    bd1, _k1 = pkg.make_polytopes(a)
    bd2, _k2 = pkg.make_polytopes(b)
    n = len(bd1)
    d_bd1, d_bd2, d_c1, d_c2, d_simp, d_dist = eng.allocate_and_copy_device_arrays(bd1, bd2)
    d_w1, d_w2, d_nrm = eng.allocate_epa_device_arrays(n)
    try:
        eng.compute_minimum_distance_device(n, d_bd1, d_bd2, d_simp, d_dist)
        simp, dist = eng.copy_results_from_device(n, d_simp, d_dist)
        (os_, od), (es, ed, en) = _oracle(oracle_mod, dtype, a, b)
        assert np.array_equal(dist, od) and live_simplex_equal(simp, os_)
        eng.compute_epa_device(n, d_bd1, d_bd2, d_simp, d_dist, d_nrm)
        simp, dist = eng.copy_results_from_device(n, d_simp, d_dist)
        _w1, _w2, nrm = eng.copy_epa_results_from_device(n, d_w1, d_w2, d_nrm)
        assert np.array_equal(dist, ed) and live_simplex_equal(simp, es)
        # normals of pairs the reference leaves untouched are whatever the buffer held: compare written ones
        wrote = ~((ed <= np.finfo(dtype).eps) & (ed == od) & (od != 0))
        assert np.array_equal(nrm[wrote], en[wrote])
    finally:
        eng.free_epa_device_arrays(d_w1, d_w2, d_nrm)
        eng.free_device_arrays(d_bd1, d_bd2, d_c1, d_c2, d_simp, d_dist)
    torch.cuda.synchronize()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("nverts,spread,n", [(64, 10.0, 30000), (32, 1.0, 30000), (8, 10.0, 20000), (16, 2.0, 20000),
                                             (128, 4.0, 4000), (256, 6.0, 2000), (12, 1.0, 9000), (20, 3.0, 9000),
                                             (30, 2.0, 5000)])
def test_uniform_device_fast_path(pkg, oracle_mod, dtype, nverts, spread, n):
    """gjk_uniform_device / epa_uniform_device on torch-owned device memory (what bench.py times)."""
    import torch
    a, b = pkg.workloads.random_pairs(n, nverts, spread, seed=77, dtype=dtype)
    eng = pkg.Engine(dtype)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    d_a, d_b = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    d_simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device="cuda")
    d_dist = torch.zeros(n, dtype=tdt, device="cuda")
    d_nrm = torch.zeros(n, 3, dtype=tdt, device="cuda")
    eng.gjk_uniform_device(n, nverts, d_a, nverts, d_b, d_simp, d_dist)
    (os_, od), (es, ed, en) = _oracle(oracle_mod, dtype, a, b)
    simp = d_simp.cpu().numpy().view(eng.sdtype)
    assert np.array_equal(d_dist.cpu().numpy(), od)
    assert live_simplex_equal(simp, os_)
    eng.epa_uniform_device(n, nverts, d_a, nverts, d_b, d_simp, d_dist, d_nrm)
    simp = d_simp.cpu().numpy().view(eng.sdtype)
    assert np.array_equal(d_dist.cpu().numpy(), ed)
    assert np.array_equal(d_nrm.cpu().numpy(), en)
    assert live_simplex_equal(simp, es)
    eng.set_stream(0)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_uniform_device_different_vertex_counts(pkg, oracle_mod, dtype):
    import torch
    n = 8000
    a = pkg.workloads.random_polytopes(n, 64, 6.0, 1, dtype, stream=1)
    b = pkg.workloads.random_polytopes(n, 16, 6.0, 1, dtype, stream=2)
    eng = pkg.Engine(dtype)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    d_a, d_b = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    d_simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device="cuda")
    d_dist = torch.zeros(n, dtype=tdt, device="cuda")
    eng.gjk_uniform_device(n, 64, d_a, 16, d_b, d_simp, d_dist)
    torch.cuda.synchronize()
    os_, od = oracle_mod.Oracle("port", dtype).gjk(a, b)
    assert np.array_equal(d_dist.cpu().numpy(), od)
    assert live_simplex_equal(d_simp.cpu().numpy().view(eng.sdtype), os_)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_indexed_api(pkg, oracle_mod, dtype):
    """compute_minimum_distance_indexed / compute_epa_indexed / compute_gjk_epa_indexed + the device-level indexed
    calls; pool interleaved as in reference examples/main.cpp:409-415 plus random re-use of pool entries."""
    rng = np.random.default_rng(9)
    counts = rng.integers(4, 40, size=600)
    pool = [pkg.workloads.random_polytopes(1, int(c), 5.0, 300 + i, dtype)[0] for i, c in enumerate(counts)]
    pairs = rng.integers(0, 600, size=(9000, 2)).astype(np.int32)
    off = np.concatenate([[0], np.cumsum(counts)])
    flat = np.concatenate(pool)
    orc = oracle_mod.Oracle("port", dtype)
    gs, gd, _ = orc.gjk_epa_indexed(flat, pairs, off, do_epa=False)
    es, ed, en = orc.gjk_epa_indexed(flat, pairs, off)
    eng = pkg.Engine(dtype)
    desc, _keep = pkg.make_polytopes(pool, dtype)
    s, d = eng.compute_minimum_distance_indexed(desc, pairs)
    assert np.array_equal(d, gd) and live_simplex_equal(s, gs)
    s2, d2, n2 = eng.compute_epa_indexed(desc, pairs, s.copy(), d.copy())
    assert np.array_equal(d2, ed) and np.array_equal(n2, en) and live_simplex_equal(s2, es)
    s3, d3, n3 = eng.compute_gjk_epa_indexed(desc, pairs)
    assert np.array_equal(d3, ed) and np.array_equal(n3, en) and live_simplex_equal(s3, es)
    # device-level
    dp, dc, dpairs, dsimp, ddist, dnrm = eng.allocate_indexed_device(desc, len(pairs))
    try:
        eng.upload_pairs_device(pairs, dpairs)
        eng.compute_minimum_distance_indexed_device(len(pairs), dp, dpairs, dsimp, ddist)
        eng.compute_epa_indexed_device(len(pairs), dp, dpairs, dsimp, ddist, dnrm)
        s4, d4 = eng.copy_results_from_device(len(pairs), dsimp, ddist)
        assert np.array_equal(d4, ed) and live_simplex_equal(s4, es)
    finally:
        eng.free_indexed_device(dp, dc, dpairs, dsimp, ddist, dnrm)


def test_readme_spelling_with_witness_arrays(pkg, oracle_mod):
    dtype = np.float32
    a, b = pkg.workloads.random_pairs(5000, 16, 1.0, seed=2, dtype=dtype)
    eng = pkg.Engine(dtype)
    bd1, _k1 = pkg.make_polytopes(a)
    bd2, _k2 = pkg.make_polytopes(b)
    simp, dist, w1, w2, nrm = eng.compute_collision_information_witness(bd1, bd2)
    (_os, _od), (es, ed, en) = _oracle(oracle_mod, dtype, a, b)
    assert np.array_equal(dist, ed) and np.array_equal(nrm, en)
    assert np.array_equal(w1, es["witnesses"][:, 0]) and np.array_equal(w2, es["witnesses"][:, 1])


def test_full_size_properties(pkg):
    """BASELINE config 2 at full size (1 Mi x 64 verts fp32): size-independent properties.
    (i) swapping the bodies leaves the distance unchanged up to rounding and swaps the witnesses;
    (ii) translating both bodies leaves verdicts unchanged; (iii) distance == |w1 - w2| for separated pairs."""
    import torch
    n, nv = 1 << 20, 64
    dtype = np.float32
    a, b = pkg.workloads.random_pairs(n, nv, 10.0, seed=12345, dtype=dtype)
    eng = pkg.Engine(dtype)
    d_a, d_b = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = []
    for x, y in ((d_a, d_b), (d_b, d_a)):
        d_simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device="cuda")
        d_dist = torch.zeros(n, dtype=torch.float32, device="cuda")
        eng.gjk_uniform_device(n, nv, x, nv, y, d_simp, d_dist)
        torch.cuda.synchronize()
        out.append((d_simp.cpu().numpy().view(eng.sdtype), d_dist.cpu().numpy()))
    (s_ab, d_ab), (s_ba, d_ba) = out
    eps = np.finfo(dtype).eps
    sep = d_ab > 1e-3
    assert 0.93 < sep.mean() < 0.97
    np.testing.assert_allclose(d_ab[sep], d_ba[sep], rtol=2e-4)
    assert np.mean((d_ab <= eps) == (d_ba <= eps)) > 0.9999
    gap = np.linalg.norm(s_ab["witnesses"][:, 0] - s_ab["witnesses"][:, 1], axis=1)
    np.testing.assert_allclose(gap[sep], d_ab[sep], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(s_ab["witnesses"][sep, 0], s_ba["witnesses"][sep, 1], rtol=0, atol=2e-2)

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


def _raw(ptr, nbytes):
    """torch uint8 view of a raw device pointer"""
    import torch

    class _P:
        def __init__(self):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3}

    return torch.as_tensor(_P(), device="cuda")


@pytest.mark.parametrize("what", ["repoint", "numpoints", "stale_address"])
def test_descriptors_edited_after_upload_are_honoured(pkg, oracle_mod, what):
    """The reference hands the device descriptor arrays to the caller, who may edit them.  The library remembers the
    layout it uploaded only as a hint and re-validates it on the device at every call, so the results must follow the
    LIVE descriptors: (a) d_bd1[i].coord re-pointed at another polytope, (b) numpoints lowered, (c) a registered array
    freed behind the library's back and its address re-used by a different upload."""
    import torch
    dtype = np.float32
    n, nv = 40000, 64
    a, b = pkg.workloads.random_pairs(n, nv, 8.0, seed=909, dtype=dtype)
    eng = pkg.Engine(dtype)
    orc = oracle_mod.Oracle("port", dtype)
    bd1, _k1 = pkg.make_polytopes(a)
    bd2, _k2 = pkg.make_polytopes(b)
    h = eng.allocate_and_copy_device_arrays(bd1, bd2)
    d_bd1, d_bd2, d_c1, d_c2, d_simp, d_dist = h
    try:
        desc = _raw(d_bd1, n * eng.pdtype.itemsize).cpu().numpy().view(eng.pdtype).copy()
        if what == "repoint":  # pair 7 now uses polytope 11 of the same blob as body 1
            desc["coord"][7] = desc["coord"][11]
            _raw(d_bd1, n * eng.pdtype.itemsize).copy_(torch.from_numpy(desc.view(np.uint8)))
            a2 = a.copy()
            a2[7] = a[11]
            want = orc.gjk(a2, b, nthreads=8)
        elif what == "numpoints":  # pair 5's body 1 shrinks to its first 9 vertices
            desc["numpoints"][5] = 9
            _raw(d_bd1, n * eng.pdtype.itemsize).copy_(torch.from_numpy(desc.view(np.uint8)))
            off1 = np.concatenate([[0], np.cumsum(desc["numpoints"])]).astype(np.int64)
            flat = np.concatenate([a[i, : desc["numpoints"][i]] for i in range(n)])
            want = orc.gjk(flat, b.reshape(-1, 3), off1, np.arange(n + 1, dtype=np.int64) * nv, nthreads=8)
        else:  # overwrite the whole descriptor array with descriptors of a reversed batch (as a recycled address would hold)
            desc["coord"] = desc["coord"][::-1].copy()
            _raw(d_bd1, n * eng.pdtype.itemsize).copy_(torch.from_numpy(desc.view(np.uint8)))
            want = orc.gjk(a[::-1].copy(), b, nthreads=8)
        torch.cuda.synchronize()
        eng.compute_minimum_distance_device(n, d_bd1, d_bd2, d_simp, d_dist)
        simp, dist = eng.copy_results_from_device(n, d_simp, d_dist)
        assert np.array_equal(dist, want[1]) and live_simplex_equal(simp, want[0])
    finally:
        eng.free_device_arrays(*h)


def test_two_streams_in_flight_from_one_thread(pkg, oracle_mod):
    """ogjk_set_stream + ogjk_set_sync(0): two fused GJK+EPA calls of one thread in flight on two streams must not
    share tickets / EPA queues (scratch is keyed by (device, stream))."""
    import torch
    dtype = np.float32
    n, nv = 60000, 32
    eng = pkg.Engine(dtype)
    orc = oracle_mod.Oracle("port", dtype)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    jobs = []
    for k, st in enumerate(streams):
        a, b = pkg.workloads.random_pairs(n, nv, 1.5, seed=300 + k, dtype=dtype)
        d_a, d_b = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        d_simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device="cuda")
        d_dist = torch.zeros(n, dtype=torch.float32, device="cuda")
        d_nrm = torch.zeros(n, 3, dtype=torch.float32, device="cuda")
        jobs.append((a, b, d_a, d_b, d_simp, d_dist, d_nrm))
    torch.cuda.synchronize()
    eng.set_sync(False)
    try:
        for _rep in range(3):
            for st, (a, b, d_a, d_b, d_simp, d_dist, d_nrm) in zip(streams, jobs):
                eng.set_stream(st.cuda_stream)
                eng.gjk_epa_uniform_device(n, nv, d_a, nv, d_b, d_simp, d_dist, d_nrm)
        torch.cuda.synchronize()
    finally:
        eng.set_sync(True)
        eng.set_stream(0)
    for a, b, _da, _db, d_simp, d_dist, d_nrm in jobs:
        s, d = orc.gjk(a, b, nthreads=8)
        es, ed, en = orc.epa(a, b, s, d, nthreads=8)
        assert np.array_equal(d_dist.cpu().numpy(), ed) and np.array_equal(d_nrm.cpu().numpy(), en)
        assert live_simplex_equal(d_simp.cpu().numpy().view(eng.sdtype), es)


def test_host_api_fans_out_over_selected_devices(pkg, oracle_mod):
    """ogjk_set_devices: the host-pointer calls slice the pair range over the selected devices (one host thread per
    device, results written at the slice offsets).  With one GPU visible the same code path runs with the device listed
    once; with >= 2 it really fans out.  Dense and indexed entry points, bit-exact."""
    import ctypes
    dtype = np.float32
    lib = pkg.load_library()
    ndev = lib.ogjk_device_count()
    use = min(ndev, 4) if ndev > 1 else 2  # one GPU: list it twice, so the slicing + worker path still runs
    eng = pkg.Engine(dtype)
    orc = oracle_mod.Oracle("port", dtype)
    n, nv = 300000, 32
    a, b = pkg.workloads.random_pairs(n, nv, 2.0, seed=515, dtype=dtype)
    bd1, _k1 = pkg.make_polytopes(a)
    bd2, _k2 = pkg.make_polytopes(b)
    s, d = orc.gjk(a, b, nthreads=8)
    es, ed, en = orc.epa(a, b, s, d, nthreads=8)
    pool, pairs = pkg.workloads.broadphase_pool(3000, 32, 300000, seed=77)
    ps, pd, pn = orc.gjk_epa_indexed(pool, pairs, nthreads=8)
    desc, _keep = pkg.make_polytopes(pool)
    devs = (ctypes.c_int * use)(*[i % ndev for i in range(use)])
    assert lib.ogjk_set_devices(ctypes.c_int(use), devs) == 0
    try:
        gs, gd, gn = eng.compute_gjk_epa(bd1, bd2)
        assert np.array_equal(gd, ed) and np.array_equal(gn, en) and live_simplex_equal(gs, es)
        gs, gd = eng.compute_minimum_distance(bd1, bd2)
        assert np.array_equal(gd, d) and live_simplex_equal(gs, s)
        qs, qd, qn = eng.compute_gjk_epa_indexed(desc, pairs)
        assert np.array_equal(qd, pd) and np.array_equal(qn, pn) and live_simplex_equal(qs, ps)
    finally:
        assert lib.ogjk_set_devices(ctypes.c_int(0), None) == 0
    assert lib.ogjk_release_cached_buffers() == 0
    gs, gd, gn = eng.compute_gjk_epa(bd1, bd2)  # buffers come back after a release
    assert np.array_equal(gd, ed) and np.array_equal(gn, en)

"""World transform + descriptor upkeep and the whole per-frame device pipeline of the
reference's caller: transform -> broad phase -> indexed GJK -> indexed EPA, against the CPU oracles."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT, live_simplex_equal


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "oracle", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _bodies(n, seed, box=9.0):
    rng = np.random.default_rng(seed)
    pos = np.zeros((n, 4), np.float32)
    pos[:, :3] = rng.uniform(-box, box, (n, 3))
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    scales = rng.uniform(0.3, 1.4, (n, 3)).astype(np.float32)
    return pos, q.astype(np.float32), scales


def test_oracle_rotation_is_a_rotation():
    tr = _load("transform_oracle")
    pos, q, sc = _bodies(200, 1)
    v = np.random.default_rng(2).standard_normal((200, 16, 3)).astype(np.float32)
    one = np.ones_like(sc)
    w = tr.transform_uniform(np.zeros_like(pos), q, one, v)
    np.testing.assert_allclose(np.linalg.norm(w, axis=2), np.linalg.norm(v, axis=2), rtol=2e-5)
    ident = np.tile(np.array([0, 0, 0, 1], np.float32), (200, 1))
    assert np.array_equal(tr.transform_uniform(pos, ident, one, v), (v + pos[:, None, :3]).astype(np.float32))


@pytest.mark.gpu
def test_device_transform_matches_oracle(pkg):
    import torch
    tr = _load("transform_oracle")
    n, nv = 4000, 32
    pos, q, sc = _bodies(n, 3)
    local = pkg.workloads.unit_sphere_hulls(n, nv, 9, np.float32)
    eng = pkg.Engine(np.float32)
    d_world = torch.zeros(n, nv, 3, dtype=torch.float32, device="cuda")
    eng.transform_to_world_device(n, torch.from_numpy(pos).cuda(), torch.from_numpy(q).cuda(), torch.from_numpy(sc).cuda(),
                                  torch.from_numpy(local).cuda(), d_world, uniform_count=nv)
    torch.cuda.synchronize()
    assert np.array_equal(d_world.cpu().numpy(), tr.transform_uniform(pos, q, sc, local))
    # ragged sub-meshes, several per body (the reference's layout: offsets, counts, owning body)
    rng = np.random.default_rng(4)
    counts = rng.integers(4, 40, 900).astype(np.int32)
    offsets = (np.cumsum(counts) - counts).astype(np.int32)
    sub_body = rng.integers(0, n, 900).astype(np.int32)
    flat = rng.standard_normal((int(counts.sum()), 3)).astype(np.float32)
    d_out = torch.zeros(flat.shape, dtype=torch.float32, device="cuda")
    eng.transform_to_world_device(900, torch.from_numpy(pos).cuda(), torch.from_numpy(q).cuda(), torch.from_numpy(sc).cuda(),
                                  torch.from_numpy(flat).cuda(), d_out, torch.from_numpy(offsets).cuda(),
                                  torch.from_numpy(counts).cuda(), torch.from_numpy(sub_body).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), tr.transform_ragged(pos, q, sc, flat, offsets, counts, sub_body))


@pytest.mark.gpu
def test_device_frame_pipeline(pkg, oracle_mod):
    """one 'frame' of the reference's caller without the physics: local hulls + poses -> world pool (transform) ->
    descriptors (init_polytopes) -> candidate pairs (broad phase) -> GJK -> EPA, everything device resident"""
    import ctypes
    import torch
    tr, bp = _load("transform_oracle"), _load("broadphase_oracle")
    n, nv = 5000, 32
    pos, q, sc3 = _bodies(n, 21)
    s = sc3[:, :1]
    sc = np.repeat(s, 3, 1).astype(np.float32)                      # uniform scale: bounding radius = scale
    local = pkg.workloads.unit_sphere_hulls(n, nv, 13, np.float32)
    spheres = pos.copy()
    spheres[:, 3] = s[:, 0] * np.float32(1.0001)
    world = tr.transform_uniform(pos, q, sc, local)
    cell, boundary, grid = 2.9, 10.0, 7
    want_pairs = bp.pairs(spheres, cell, boundary, grid)
    assert want_pairs.shape[0] >= 32768
    eng = pkg.Engine(np.float32)
    d_world = torch.zeros(n, nv, 3, dtype=torch.float32, device="cuda")
    d_desc = torch.zeros(n * eng.pdtype.itemsize, dtype=torch.uint8, device="cuda")
    cap = want_pairs.shape[0]
    d_pairs = torch.zeros(cap, 2, dtype=torch.int32, device="cuda")
    d_simp = torch.zeros(cap * eng.sdtype.itemsize, dtype=torch.uint8, device="cuda")
    d_dist = torch.zeros(cap, dtype=torch.float32, device="cuda")
    d_nrm = torch.zeros(cap, 3, dtype=torch.float32, device="cuda")
    eng.transform_to_world_device(n, torch.from_numpy(pos).cuda(), torch.from_numpy(q).cuda(), torch.from_numpy(sc).cuda(),
                                  torch.from_numpy(local).cuda(), d_world, uniform_count=nv)
    eng.init_polytopes_device(d_desc, d_world, n, uniform_count=nv)
    try:
        total = eng.broadphase_pairs_device(n, torch.from_numpy(spheres).cuda(), cell, boundary, grid, d_pairs, cap)
        assert total == cap
        eng.compute_minimum_distance_indexed_device(total, d_desc, d_pairs, d_simp, d_dist)
        eng.compute_epa_indexed_device(total, d_desc, d_pairs, d_simp, d_dist, d_nrm)
        torch.cuda.synchronize()
    finally:
        eng.release_pool(d_desc)
    assert np.array_equal(d_world.cpu().numpy(), world)
    got_pairs = d_pairs.cpu().numpy()
    off = np.arange(n + 1) * nv
    es, ed, en = oracle_mod.Oracle("port", np.float32).gjk_epa_indexed(world.reshape(-1, 3), got_pairs, off, nthreads=8)
    assert np.array_equal(d_dist.cpu().numpy(), ed)
    assert np.array_equal(d_nrm.cpu().numpy(), en)
    assert live_simplex_equal(d_simp.cpu().numpy().view(eng.sdtype), es)
    assert np.array_equal(got_pairs[np.lexsort((got_pairs[:, 1], got_pairs[:, 0]))], want_pairs)

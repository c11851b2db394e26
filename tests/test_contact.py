"""Contact response: the consumer of distances / witnesses / contact normals, against the
numpy restatement of the reference's collision_response_kernel (oracle/contact_oracle.py)."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "oracle", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _state(nb, seed, box=4.0):
    rng = np.random.default_rng(seed)
    pos = np.zeros((nb, 4), np.float32)
    pos[:, :3] = rng.uniform(-box, box, (nb, 3))
    pos[:, 3] = rng.uniform(0.3, 1.5, nb)                      # w = bounding radius, never touched
    vel = np.zeros((nb, 4), np.float32)
    vel[:, :3] = rng.normal(0, 3.0, (nb, 3))
    vel[:, 3] = rng.uniform(0.5, 4.0, nb)                      # w = mass (reference: inv_m = 1 / vel.w)
    ang = np.zeros((nb, 4), np.float32)
    ang[:, :3] = rng.normal(0, 1.0, (nb, 3))
    q = rng.standard_normal((nb, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    inv_i = rng.uniform(0.2, 3.0, (nb, 3)).astype(np.float32)
    return pos, vel, ang, q.astype(np.float32), inv_i


def _contacts(npairs, nb, seed, dtype=np.float32, hub=None):
    """synthetic GJK/EPA outputs: a mix of separated (> eps), touching (0) and penetrating (< 0) pairs, some with a
    degenerate (zero) normal; `hub` makes one body take part in every third pair"""
    rng = np.random.default_rng(seed)
    pairs = rng.integers(0, nb, (npairs, 2)).astype(np.int32)
    same = pairs[:, 0] == pairs[:, 1]
    pairs[same, 1] = (pairs[same, 0] + 1) % nb
    if hub is not None:
        pairs[::3, rng.integers(0, 2)] = hub
        pairs[pairs[:, 0] == pairs[:, 1], 1] = (hub + 1) % nb
    kind = rng.integers(0, 4, npairs)
    dist = np.where(kind == 0, rng.uniform(0.01, 2.0, npairs), np.where(kind == 1, 0.0, -rng.uniform(1e-4, 0.4, npairs)))
    nrm = rng.standard_normal((npairs, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm[rng.random(npairs) < 0.02] = 0.0
    wit = rng.uniform(-4.0, 4.0, (npairs, 2, 3))
    return pairs, dist.astype(dtype), wit.astype(dtype), nrm.astype(dtype)


def test_oracle_accumulation_order_and_physics():
    co = _load("contact_oracle")
    nb, npairs = 40, 600
    pos, vel, ang, q, inv_i = _state(nb, 1)
    pairs, dist, wit, nrm = _contacts(npairs, nb, 2, hub=7)
    a = co.contact_response(pairs, dist, wit, nrm, None, pos, vel, ang, q, inv_i)
    b = co.contact_response_loop(pairs, dist, wit, nrm, None, pos, vel, ang, q, inv_i)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert np.array_equal(a[0][:, 3], pos[:, 3]) and np.array_equal(a[1][:, 3], vel[:, 3])
    # one head-on pair: equal masses, no spin, contact on the line of centres -> velocities exchange with e = 0.7
    p1 = np.array([[0, 0, 0, 1], [1.5, 0, 0, 1]], np.float32)
    v1 = np.array([[3, 0, 0, 2], [-3, 0, 0, 2]], np.float32)
    z = np.zeros((2, 4), np.float32)
    ident = np.tile(np.array([0, 0, 0, 1], np.float32), (2, 1))
    one = np.ones((2, 3), np.float32)
    w1 = np.array([[[1.0, 0, 0], [0.5, 0, 0]]], np.float32)
    po, vo, ao = co.contact_response(np.array([[0, 1]], np.int32), np.array([-0.5], np.float32), w1,
                                     np.array([[1.0, 0, 0]], np.float32), None, p1, v1, z, ident, one)
    np.testing.assert_allclose(vo[:, 0], [3 - 1.7 * 3, -3 + 1.7 * 3], rtol=1e-6)      # j = (1+e) * 6 / (1/2 + 1/2)
    np.testing.assert_allclose(po[:, 0], [0 - 0.2 * 0.5 * 0.5, 1.5 + 0.2 * 0.5 * 0.5], rtol=1e-6)
    assert np.array_equal(ao, z)
    # momentum: sum m * dv = 0 over every pair
    t = co.pair_terms(pairs, dist, wit, nrm, None, pos, vel, ang, q, inv_i)
    m = t["has_vel"]
    ma, mb = vel[t["idA"], 3][m, None], vel[t["idB"], 3][m, None]
    np.testing.assert_allclose(ma * t["dvel_A"][m] + mb * t["dvel_B"][m], 0, atol=2e-4)
    # separated pairs and pairs moving apart contribute nothing
    assert not t["has_pos"][dist > 0].any() and not t["has_vel"][dist > 0].any()


def _run_device(pkg, dtype, pairs, dist, wit, nrm, sub_body, pos, vel, ang, q, inv_i, **prm):
    import torch
    eng = pkg.Engine(dtype)
    n = len(pairs)
    simp = np.zeros(max(n, 1), eng.sdtype)
    simp["witnesses"][:n] = wit
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_pos, d_vel, d_ang = dev(pos), dev(vel), dev(ang)
    d_vel2 = torch.full_like(d_vel, float("nan"))
    d_ang2 = torch.full_like(d_ang, float("nan"))
    eng.contact_response_device(n, dev(pairs) if n else None, dev(dist) if n else None,
                                dev(simp.view(np.uint8)) if n else None, dev(nrm) if n else None, len(pos), d_pos,
                                d_vel, d_vel2, d_ang, d_ang2, dev(q), dev(inv_i),
                                d_sub_mesh_body=None if sub_body is None else dev(sub_body), **prm)
    torch.cuda.synchronize()
    assert np.array_equal(d_vel.cpu().numpy(), vel) and np.array_equal(d_ang.cpu().numpy(), ang)  # ping untouched
    return d_pos.cpu().numpy(), d_vel2.cpu().numpy(), d_ang2.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("nb,npairs,hub", [(300, 20000, None), (64, 50000, 5), (5000, 7, None), (33, 1, None)])
def test_device_contact_response_matches_oracle(pkg, dtype, nb, npairs, hub):
    co = _load("contact_oracle")
    pos, vel, ang, q, inv_i = _state(nb, 11 + nb)
    pairs, dist, wit, nrm = _contacts(npairs, nb, 5 + npairs, dtype, hub)
    prm = dict(epsilon=1e-6, restitution=0.7, restitution_threshold=2.0, baumgarte_beta=0.2)
    want = co.contact_response(pairs, dist, wit, nrm, None, pos, vel, ang, q, inv_i, **prm)
    got = _run_device(pkg, dtype, pairs, dist, wit, nrm, None, pos, vel, ang, q, inv_i, **prm)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    # twice in a row gives the same bits (no dependence on scheduling)
    again = _run_device(pkg, dtype, pairs, dist, wit, nrm, None, pos, vel, ang, q, inv_i, **prm)
    for g, w in zip(again, got):
        assert np.array_equal(g, w)


@pytest.mark.gpu
def test_device_contact_response_sub_meshes_and_edges(pkg):
    """pairs index sub-meshes; sub_mesh_body maps them to bodies, invalid owners (-1, >= bodies) are skipped; no pairs
    at all still performs the ping -> pong copy"""
    co = _load("contact_oracle")
    nb, nsub, npairs = 120, 700, 30000
    rng = np.random.default_rng(3)
    sub_body = rng.integers(0, nb, nsub).astype(np.int32)
    sub_body[rng.random(nsub) < 0.03] = -1
    sub_body[rng.random(nsub) < 0.03] = nb + 4
    pos, vel, ang, q, inv_i = _state(nb, 8)
    pairs, dist, wit, nrm = _contacts(npairs, nsub, 9)
    want = co.contact_response(pairs, dist, wit, nrm, sub_body, pos, vel, ang, q, inv_i, epsilon=0.0)
    got = _run_device(pkg, np.float32, pairs, dist, wit, nrm, sub_body, pos, vel, ang, q, inv_i, epsilon=0.0)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    assert not np.array_equal(got[1], vel)
    got0 = _run_device(pkg, np.float32, pairs[:0], dist[:0], wit[:0], nrm[:0], None, pos, vel, ang, q, inv_i)
    assert np.array_equal(got0[0], pos) and np.array_equal(got0[1], vel) and np.array_equal(got0[2], ang)


@pytest.mark.gpu
def test_device_frame_with_contact_response(pkg, oracle_mod):
    """broad phase -> indexed GJK -> EPA -> contact response on the device against the CPU oracles end to end"""
    import torch
    tr, bp, co = _load("transform_oracle"), _load("broadphase_oracle"), _load("contact_oracle")
    n, nv = 3000, 32
    rng = np.random.default_rng(77)
    pos, vel, ang, q, inv_i = _state(n, 70, box=7.0)
    s = rng.uniform(0.4, 1.2, (n, 1)).astype(np.float32)
    sc = np.repeat(s, 3, 1).astype(np.float32)
    pos[:, 3] = s[:, 0] * np.float32(1.0001)
    local = pkg.workloads.unit_sphere_hulls(n, nv, 5, np.float32)
    world = tr.transform_uniform(pos, q, sc, local)
    cell, boundary, grid = 2.5, 8.0, 7
    want_pairs = bp.pairs(pos, cell, boundary, grid)
    cap = want_pairs.shape[0]
    eng = pkg.Engine(np.float32)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_world = dev(world)
    d_desc = torch.zeros(n * eng.pdtype.itemsize, dtype=torch.uint8, device="cuda")
    d_pairs = torch.zeros(cap, 2, dtype=torch.int32, device="cuda")
    d_simp = torch.zeros(cap * eng.sdtype.itemsize, dtype=torch.uint8, device="cuda")
    d_dist = torch.zeros(cap, dtype=torch.float32, device="cuda")
    d_nrm = torch.zeros(cap, 3, dtype=torch.float32, device="cuda")
    d_pos, d_vel, d_ang = dev(pos), dev(vel), dev(ang)
    d_vel2, d_ang2 = torch.zeros_like(d_vel), torch.zeros_like(d_ang)
    eng.init_polytopes_device(d_desc, d_world, n, uniform_count=nv)
    try:
        total = eng.broadphase_pairs_device(n, d_pos, cell, boundary, grid, d_pairs, cap)
        assert total == cap
        eng.compute_minimum_distance_indexed_device(total, d_desc, d_pairs, d_simp, d_dist)
        eng.compute_epa_indexed_device(total, d_desc, d_pairs, d_simp, d_dist, d_nrm)
        eng.contact_response_device(total, d_pairs, d_dist, d_simp, d_nrm, n, d_pos, d_vel, d_vel2, d_ang, d_ang2,
                                    dev(q), dev(inv_i), epsilon=1e-6)
        torch.cuda.synchronize()
    finally:
        eng.release_pool(d_desc)
    got_pairs = d_pairs.cpu().numpy()
    off = np.arange(n + 1) * nv
    es, ed, en = oracle_mod.Oracle("port", np.float32).gjk_epa_indexed(world.reshape(-1, 3), got_pairs, off, nthreads=8)
    assert np.array_equal(d_dist.cpu().numpy(), ed) and np.array_equal(d_nrm.cpu().numpy(), en)
    want = co.contact_response(got_pairs, ed, es["witnesses"], en, None, pos, vel, ang, q, inv_i, epsilon=1e-6)
    assert (ed <= 1e-6).sum() > 100
    for g, w in zip((d_pos.cpu().numpy(), d_vel2.cpu().numpy(), d_ang2.cpu().numpy()), want):
        assert np.array_equal(g, w)

"""Sanity of the decision tables (gjk_tables.h) through the host harness build: exercised indirectly by
test_core_host.py; here we check the workload generators are deterministic (the GPU box must see the same bytes)."""
import hashlib

import numpy as np


def test_workloads_are_deterministic(pkg):
    a1, b1 = pkg.workloads.random_pairs(1000, 64, 10.0, seed=12345)
    a2, b2 = pkg.workloads.random_pairs(1000, 64, 10.0, seed=12345)
    assert np.array_equal(a1, a2) and np.array_equal(b1, b2)
    assert not np.array_equal(a1, b1)
    # chunking must not change the stream: a prefix of a larger batch equals the smaller batch only within a chunk
    c, _ = pkg.workloads.random_pairs(10, 8, 1.0, seed=7)
    d, _ = pkg.workloads.random_pairs(10, 8, 1.0, seed=7)
    assert hashlib.sha1(c.tobytes()).hexdigest() == hashlib.sha1(d.tobytes()).hexdigest()


def test_broadphase_pool_pairs_overlap(pkg):
    pool, pairs = pkg.workloads.broadphase_pool(400, 16, 3000, seed=3)
    assert pool.shape == (400, 16, 3) and pairs.shape[1] == 2 and len(pairs) > 100
    assert np.all(pairs[:, 0] < pairs[:, 1])
    c = pool.mean(axis=1)
    r = np.linalg.norm(pool - c[:, None], axis=2).max(axis=1)
    d = np.linalg.norm(c[pairs[:, 0]] - c[pairs[:, 1]], axis=1)
    assert np.all(d < 1.05 * (r[pairs[:, 0]] + r[pairs[:, 1]]))

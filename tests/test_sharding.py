"""N>1 host logic on CPU: two gloo ranks shard a pair batch, each solves its slice (with the CPU oracle standing in
for the device), slices are gathered and must reproduce the single-process result; timing reduction = MAX."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_range(pkg):
    from itertools import chain
    sb = pkg.sharding.shard_bounds
    for n in (0, 1, 7, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            parts = [sb(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
            assert list(chain.from_iterable(range(lo, hi) for lo, hi in parts)) == list(range(n))


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, {root!r})
    from _pkgpath import load_package, load_oracle
    pkg, om = load_package(), load_oracle()
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 3001
    a, b = pkg.workloads.random_pairs(n, 16, 2.0, seed=5)
    lo, hi = pkg.sharding.shard_bounds(n, rank, world)
    orc = om.Oracle("port", np.float32)
    s, d = orc.gjk(a[lo:hi], b[lo:hi])
    s, d, nrm = orc.epa(a[lo:hi], b[lo:hi], s, d)
    full_d = pkg.sharding.gather_slices(torch.from_numpy(d), n, rank, world).numpy()
    full_n = pkg.sharding.gather_slices(torch.from_numpy(nrm), n, rank, world).numpy()
    t = pkg.sharding.max_over_ranks([1.0 + rank, 5.0 - rank])
    if rank == 0:
        s0, d0 = orc.gjk(a, b)
        s0, d0, n0 = orc.epa(a, b, s0, d0)
        assert np.array_equal(full_d, d0) and np.array_equal(full_n, n0)
        assert t == [float(world), 5.0], t
        print("SHARD_OK")
    dist.destroy_process_group()
""")


def test_two_rank_gloo_sharding(tmp_path, oracle_mod):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SHARD_OK" in out.stdout

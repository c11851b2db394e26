"""Soak test of the persistent slot kernels: many launches over varied sizes / vertex counts / overlap, every result
compared with the CPU oracle (bit-exact).  Usage: python scripts/stress_slots.py [ROUNDS]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package, load_oracle
pkg = load_package(); orc = load_oracle().Oracle('port', np.float32)
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(2026)
eng = pkg.Engine(np.float32); eng.set_device(0)
t0 = time.time(); bad = 0
for r in range(rounds):
    nv = int(rng.choice([8, 16, 24, 32, 40, 48, 64, 64, 64, 80, 96, 128, 140]))
    n = int(rng.integers(32768, 140000))
    spread = float(rng.choice([0.5, 1.0, 3.0, 10.0]))
    fused = bool(rng.integers(0, 2))
    a, b = pkg.workloads.random_pairs(n, nv, spread, seed=int(rng.integers(1 << 30)), dtype=np.float32)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device='cuda')
    dist = torch.zeros(n, dtype=torch.float32, device='cuda'); nrm = torch.zeros(n, 3, dtype=torch.float32, device='cuda')
    for rep in range(3):  # repeated launches on the same data: scheduling differs, results must not
        if fused: eng.gjk_epa_uniform_device(n, nv, da, nv, db, simp, dist, nrm)
        else: eng.gjk_uniform_device(n, nv, da, nv, db, simp, dist)
    torch.cuda.synchronize()
    s, d = orc.gjk(a, b, nthreads=16)
    if fused: s, d, nr = orc.epa(a, b, s, d, nthreads=16)
    ok = np.array_equal(dist.cpu().numpy(), d) and np.array_equal(simp.cpu().numpy().view(eng.sdtype)['witnesses'], s['witnesses'])
    if fused: ok = ok and np.array_equal(nrm.cpu().numpy(), nr)
    bad += 0 if ok else 1
    print(f"round {r:3d} n={n:6d} V={nv:3d} S={spread:4.1f} fused={int(fused)} {'ok' if ok else 'MISMATCH'}", flush=True)
print(f"done: {rounds} rounds, {bad} mismatches, {time.time()-t0:.0f}s")
sys.exit(1 if bad else 0)

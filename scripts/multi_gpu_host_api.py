"""In-library multi-GPU: ONE process, the reference-facing host-pointer API, ogjk_set_devices over the visible GPUs.
Times compute_gjk_epa (dense, pinned host arrays) and compute_gjk_epa_indexed (config 5) on 1 device and on all, and
checks the fanned-out result against the single-device one bit for bit.
Usage: python scripts/multi_gpu_host_api.py [PAIRS_DENSE] [PAIRS_CFG5]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package
pkg = load_package()
eng = pkg.Engine(np.float32)
ndev = pkg.load_library().ogjk_device_count()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4 << 20
n5 = int(sys.argv[2]) if len(sys.argv) > 2 else 16_000_000
nv = 64
out = {"devices": ndev, "dense_pairs": n, "verts": nv}
a, b = pkg.workloads.random_pairs(n, nv, 10.0, seed=12345, dtype=np.float32)
pa, pb = torch.from_numpy(a).pin_memory().numpy(), torch.from_numpy(b).pin_memory().numpy()
bd1, _k1 = pkg.make_polytopes(pa); bd2, _k2 = pkg.make_polytopes(pb)
simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8).pin_memory().numpy().view(eng.sdtype)
dist = torch.zeros(n, dtype=torch.float32).pin_memory().numpy()
nrm = torch.zeros(n, 3, dtype=torch.float32).pin_memory().numpy()
ref = None
for devs in ([0], list(range(ndev))) if ndev > 1 else ([0], [0, 0]):
    eng.set_devices(devs)
    eng.compute_gjk_epa(bd1, bd2, simp, dist, nrm)  # warm-up (allocations, per-device tables)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); eng.compute_gjk_epa(bd1, bd2, simp, dist, nrm); ts.append(time.perf_counter() - t0)
    key = f"dense_{len(devs)}dev"
    out[key] = {"ms": min(ts) * 1e3, "pairs_per_s": n / min(ts), "gb_per_s_h2d": 2 * n * nv * 12 / min(ts) / 1e9}
    if ref is None:
        ref = (simp.copy(), dist.copy(), nrm.copy())
    else:
        out["dense_fanout_bit_identical"] = bool(np.array_equal(dist, ref[1]) and np.array_equal(nrm, ref[2]) and
                                                 np.array_equal(simp["witnesses"], ref[0]["witnesses"]))
    print(key, out[key], flush=True)
del a, b, pa, pb, simp, dist, nrm, ref
pool, pairs = pkg.workloads.broadphase_pool(20000, 32, n5)
desc, _keep = pkg.make_polytopes(pool)
ref = None
for devs in ([0], list(range(ndev))) if ndev > 1 else ([0], [0, 0]):
    eng.set_devices(devs)
    n5p = pairs.shape[0]
    for kind in ("pageable", "pinned"):  # result arrays: fresh numpy arrays per call / pinned arrays allocated once
        outp = None
        if kind == "pinned":
            outp = (torch.empty(n5p * eng.sdtype.itemsize, dtype=torch.uint8, pin_memory=True).numpy().view(eng.sdtype),
                    torch.empty(n5p, dtype=torch.float32, pin_memory=True).numpy(),
                    torch.empty((n5p, 3), dtype=torch.float32, pin_memory=True).numpy())
        eng.compute_gjk_epa_indexed(desc, pairs, out=outp)
        ts = []
        for _ in range(2):
            t0 = time.perf_counter(); s, d, nr = eng.compute_gjk_epa_indexed(desc, pairs, out=outp); ts.append(time.perf_counter() - t0)
        key = f"cfg5_{len(devs)}dev" + ("" if kind == "pageable" else "_pinned_out")
        out[key] = {"ms": min(ts) * 1e3, "pairs_per_s": n5p / min(ts)}
        print(key, out[key], flush=True)
    key = f"cfg5_{len(devs)}dev"
    if ref is None:
        ref = (s.copy(), d.copy(), nr.copy())
    else:
        out["cfg5_fanout_bit_identical"] = bool(np.array_equal(d, ref[1]) and np.array_equal(nr, ref[2]) and
                                                np.array_equal(s["witnesses"], ref[0]["witnesses"]))
eng.set_devices([])
print(json.dumps(out))

"""A/B of the SoA-4 packed pool on indexed batches (BASELINE config 5: 20 000-hull pool, 32 or 64 vertices, ~4 M
broad-phase pairs): GJK alone through the indexed device entry, OGJK_POOL_PACK=0 against 1, median of 10, outputs
compared bit for bit between the two and against the oracle on the first 100 k pairs."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package, load_oracle
pkg = load_package()
eng = pkg.Engine(np.float32); eng.set_device(0); eng.set_sync(False)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
for nverts in (32, 64):
    pool, pairs = pkg.workloads.broadphase_pool(20000, nverts, 4_000_000)
    n = pairs.shape[0]
    bd, _keep = pkg.make_polytopes(pool)
    dp, dc, dpairs, dsimp, ddist, dnrm = eng.allocate_indexed_device(bd, n)
    eng.upload_pairs_device(pairs, dpairs)
    m = 100000
    off = np.arange(pool.shape[0] + 1) * nverts
    os_, od, _ = load_oracle().Oracle("port", np.float32).gjk_epa_indexed(pool.reshape(-1, 3), pairs[:m], off, do_epa=False, nthreads=8)
    first = None
    for pack in ("0", "1"):
        os.environ["OGJK_POOL_PACK"] = pack
        step = lambda: eng.compute_minimum_distance_indexed_device(n, dp, dpairs, dsimp, ddist)
        for _ in range(3): step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        s, d = eng.copy_results_from_device(n, dsimp, ddist)
        ok = np.array_equal(d[:m], od) and np.array_equal(s["witnesses"][:m], os_["witnesses"]) and np.array_equal(s["nvrtx"][:m], os_["nvrtx"])
        same = True if first is None else (np.array_equal(d, first[1]) and np.array_equal(s["witnesses"], first[0]["witnesses"]) and np.array_equal(s["nvrtx"], first[0]["nvrtx"]))
        if first is None: first = (s.copy(), d.copy())
        print(f"pool 20000 x {nverts} verts, {n} pairs, pack={pack}: {np.median(ts):.3f} ms (incl. validation + pack pass)  {n/np.median(ts)*1e3:.3e} pairs/s  "
              f"kernel='{eng.last_kernel()}' oracle_eq={ok} same_as_unpacked={same}", flush=True)
    os.environ.pop("OGJK_POOL_PACK", None)
    eng.free_indexed_device(dp, dc, dpairs, dsimp, ddist, dnrm)

"""A/B of the EPA kernel families on device-resident batches (fused GJK+EPA entry, per-stage library timing).
Usage: python scripts/ab_epa.py [modes...]   (modes: warp group small4 small8 auto)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package
pkg = load_package()
modes = sys.argv[1:] or ["warp", "group", "small4", "small8", "auto"]
eng = pkg.Engine(np.float32); eng.set_device(0); eng.set_sync(False)
eng.set_stream(torch.cuda.current_stream().cuda_stream)

def run(tag, step, n):
    for mode in modes:
        if mode == "auto": os.environ.pop("OGJK_EPA_KERNEL", None)
        else: os.environ["OGJK_EPA_KERNEL"] = mode
        for _ in range(2): step()
        torch.cuda.synchronize()
        eng.set_timing(True)
        for _ in range(5): step()
        torch.cuda.synchronize()
        g, e, c = eng.stage_times(); eng.set_timing(False)
        print(f"{tag} epa={mode}: gjk {g/c:.3f} ms  epa {e/c:.3f} ms  -> {n/((g+e)/c)*1e3:.3e} pairs/s", flush=True)
    os.environ.pop("OGJK_EPA_KERNEL", None)

for name, n, nv, spread in (("cfg3 1Mi x32 S=1", 1 << 20, 32, 1.0), ("cfg2 1Mi x64 S=10", 1 << 20, 64, 10.0),
                            ("deep 512k x64 S=2", 1 << 19, 64, 2.0), ("small 1Mi x16 S=0.5", 1 << 20, 16, 0.5)):
    a, b = pkg.workloads.random_pairs(n, nv, spread, seed=12345, dtype=np.float32)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device="cuda")
    dist = torch.zeros(n, dtype=torch.float32, device="cuda"); nrm = torch.zeros(n, 3, dtype=torch.float32, device="cuda")
    run(name, lambda: eng.gjk_epa_uniform_device(n, nv, da, nv, db, simp, dist, nrm), n)
    del da, db, simp, dist, nrm

pool, pairs = pkg.workloads.broadphase_pool(20000, 32, 4_000_000)
n = pairs.shape[0]
bd, _keep = pkg.make_polytopes(pool)
dp, dc, dpairs, dsimp, ddist, dnrm = eng.allocate_indexed_device(bd, n)
eng.upload_pairs_device(pairs, dpairs)
run(f"cfg5 {n} pairs x32 indexed", lambda: eng.gjk_epa_indexed_device(n, dp, dpairs, dsimp, ddist, dnrm), n)
eng.free_indexed_device(dp, dc, dpairs, dsimp, ddist, dnrm)

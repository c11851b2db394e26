"""A/B of the EPA group kernel's service batching (OGJK_EPA_SVC=<batch><defer>) and work area for bodies of up to 16
vertices (OGJK_EPA_AREA=small|tiny) on configs 3 and 5, 16-vertex bodies and config 2.  Variants: <svc>[:<area>].
(The warps-per-CTA, horizon and 8-lane variants timed with earlier versions of this script -- profiles/r2x_*, r2y*_ab_epa_svc.txt
-- were compile-time instantiations that have since been removed.)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package
pkg = load_package()
eng = pkg.Engine(np.float32); eng.set_device(0); eng.set_sync(False)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
variants = sys.argv[1:] or ["10", "24", "34", "24:small"]
def run(tag, step, n):
    for v in variants:
        f = v.split(":")
        os.environ["OGJK_EPA_SVC"] = f[0]
        if len(f) > 1: os.environ["OGJK_EPA_AREA"] = f[1]
        else: os.environ.pop("OGJK_EPA_AREA", None)
        for _ in range(2): step()
        torch.cuda.synchronize()
        eng.set_timing(True)
        for _ in range(5): step()
        torch.cuda.synchronize()
        g, e, c = eng.stage_times(); eng.set_timing(False)
        print(f"{tag} svc={v}: gjk {g/c:.3f} ms  epa {e/c:.3f} ms", flush=True)
    for k in ("OGJK_EPA_SVC", "OGJK_EPA_AREA"): os.environ.pop(k, None)
for name, n, nv, spread in (("cfg3 1Mi x32 S=1", 1 << 20, 32, 1.0), ("small 1Mi x16 S=0.5", 1 << 20, 16, 0.5), ("cfg2 1Mi x64 S=10", 1 << 20, 64, 10.0)):
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

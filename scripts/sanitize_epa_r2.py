"""The EPA group kernel's instantiations of the last round-2 session under compute-sanitizer memcheck: lean area (32-vertex
bodies, dense and indexed sources), 1.4 KB area (16-vertex bodies), 1.7 KB area (OGJK_EPA_AREA=small), each with its overflow pass.
    compute-sanitizer --tool memcheck python scripts/sanitize_epa_r2.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package, load_oracle
pkg = load_package()
orc = load_oracle().Oracle('port', np.float32)
eng = pkg.Engine(np.float32); eng.set_device(0)
def dense(nv, n, spread, area=None):
    if area: os.environ["OGJK_EPA_AREA"] = area
    else: os.environ.pop("OGJK_EPA_AREA", None)
    a, b = pkg.workloads.random_pairs(n, nv, spread, seed=3, dtype=np.float32)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device='cuda')
    dist = torch.zeros(n, dtype=torch.float32, device='cuda'); nrm = torch.zeros(n, 3, dtype=torch.float32, device='cuda')
    os_, od = orc.gjk(a, b, nthreads=8)
    eng.gjk_epa_uniform_device(n, nv, da, nv, db, simp, dist, nrm)
    torch.cuda.synchronize()
    es, ed, en = orc.epa(a, b, os_, od, nthreads=8)
    ok = np.array_equal(dist.cpu().numpy(), ed) and np.array_equal(nrm.cpu().numpy(), en)
    print(f"dense nv={nv} n={n} area={area or 'default'}: equal={ok}", flush=True)
dense(32, 9000, 1.0)
dense(16, 9000, 0.5)
dense(32, 9000, 1.0, "small")
dense(24, 9000, 0.7)
os.environ.pop("OGJK_EPA_AREA", None)
pool, pairs = pkg.workloads.broadphase_pool(1500, 32, 45000, seed=17)
off = np.arange(pool.shape[0] + 1) * 32
gs, gd, gn = orc.gjk_epa_indexed(pool.reshape(-1, 3), pairs, off, do_epa=True, nthreads=8)
desc, _keep = pkg.make_polytopes(pool)
s, d, nr = eng.compute_gjk_epa_indexed(desc, pairs)
print(f"indexed gjk+epa: equal={np.array_equal(d, gd) and np.array_equal(nr, gn)} pairs={len(pairs)}", flush=True)

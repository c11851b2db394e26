"""Round-2 kernels under compute-sanitizer memcheck: small forced batches of the fp16 pre-scan kernel (dense + indexed),
the SoA-4 packed pool, every EPA kernel family (small work area incl. its overflow pass), and the two-set scan."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package, load_oracle
pkg = load_package()
orc = load_oracle().Oracle('port', np.float32)
eng = pkg.Engine(np.float32); eng.set_device(0)
def dense(kernel, nv, n, spread, epa_kernel=None, epa=False):
    os.environ["OGJK_GJK_KERNEL"] = kernel
    if epa_kernel: os.environ["OGJK_EPA_KERNEL"] = epa_kernel
    else: os.environ.pop("OGJK_EPA_KERNEL", None)
    a, b = pkg.workloads.random_pairs(n, nv, spread, seed=3, dtype=np.float32)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device='cuda')
    dist = torch.zeros(n, dtype=torch.float32, device='cuda'); nrm = torch.zeros(n, 3, dtype=torch.float32, device='cuda')
    os_, od = orc.gjk(a, b, nthreads=8)
    if epa:
        eng.gjk_epa_uniform_device(n, nv, da, nv, db, simp, dist, nrm)
        es, ed, en = orc.epa(a, b, os_, od, nthreads=8)
        ok = np.array_equal(dist.cpu().numpy(), ed) and np.array_equal(nrm.cpu().numpy(), en)
    else:
        eng.gjk_uniform_device(n, nv, da, nv, db, simp, dist)
        ok = np.array_equal(dist.cpu().numpy(), od)
    torch.cuda.synchronize()
    print(f"{kernel} nv={nv} n={n} epa={epa_kernel if epa else '-'}: equal={ok} kernel='{eng.last_kernel()}'", flush=True)
dense("slots16", 64, 3000, 4.0)
dense("slots16", 32, 3000, 1.0)
dense("slots16", 48, 2000, 2.0, epa=True)
dense("slotsws", 64, 3000, 4.0)
dense("slotsws", 96, 2000, 4.0)
for ek in ("warp", "small4", "small8", "group"):
    dense("auto", 32, 9000, 1.0, epa_kernel=ek, epa=True)
dense("auto", 16, 9000, 0.5, epa_kernel="small4", epa=True)
os.environ["OGJK_GJK_KERNEL"] = "slots"
for pack in ("1", "0"):
    os.environ["OGJK_POOL_PACK"] = pack
    pool, pairs = pkg.workloads.broadphase_pool(1500, 32, 45000, seed=17)
    off = np.arange(pool.shape[0] + 1) * 32
    gs, gd, _ = orc.gjk_epa_indexed(pool.reshape(-1, 3), pairs, off, do_epa=False, nthreads=8)
    desc, _keep = pkg.make_polytopes(pool)
    s, d = eng.compute_minimum_distance_indexed(desc, pairs)
    print(f"indexed pack={pack}: equal={np.array_equal(d, gd)} kernel='{eng.last_kernel()}' pairs={len(pairs)}", flush=True)
os.environ["OGJK_GJK_KERNEL"] = "slots16"
s, d = eng.compute_minimum_distance_indexed(desc, pairs)
print(f"indexed slots16: equal={np.array_equal(d, gd)} kernel='{eng.last_kernel()}'", flush=True)

"""Small forced-kernel batches for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package, load_oracle
pkg = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
for kernel, nv, epa in (("slots", 32, False), ("slotsws", 64, False), ("slotsws", 32, False), ("auto", 32, True)):
    os.environ["OGJK_GJK_KERNEL"] = kernel
    os.environ["OGJK_EPA_KERNEL"] = "group" if epa else "warp"
    a, b = pkg.workloads.random_pairs(n, nv, 1.5, seed=3, dtype=np.float32)
    eng = pkg.Engine(np.float32); eng.set_device(0)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device='cuda')
    dist = torch.zeros(n, dtype=torch.float32, device='cuda'); nrm = torch.zeros(n, 3, dtype=torch.float32, device='cuda')
    eng.gjk_uniform_device(n, nv, da, nv, db, simp, dist)
    if epa:
        bd1, _k1 = pkg.make_polytopes(a); bd2, _k2 = pkg.make_polytopes(b)
        n2 = 9000
        a2, b2 = pkg.workloads.random_pairs(n2, nv, 1.0, seed=4, dtype=np.float32)
        bd1, _k1 = pkg.make_polytopes(a2); bd2, _k2 = pkg.make_polytopes(b2)
        eng.compute_gjk_epa(bd1, bd2)
    torch.cuda.synchronize()
    orc = load_oracle().Oracle('port', np.float32)
    os_, od = orc.gjk(a, b, nthreads=8)
    print(kernel, nv, "dist_eq", np.array_equal(dist.cpu().numpy(), od), flush=True)

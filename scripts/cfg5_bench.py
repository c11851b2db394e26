"""Config 5 (BASELINE configs[4]): pool of 20,000 32-vertex hulls, broad-phase candidate pairs, indexed GJK + EPA,
device resident.  Usage: python scripts/cfg5_bench.py [NPAIRS] [NPOLY]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package
pkg = load_package()
npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
npoly = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
t0 = time.time()
pool, pairs, spheres, edge = pkg.workloads.broadphase_pool(npoly, 32, npairs, seed=7, return_spheres=True)
n = pairs.shape[0]
print(f"generated {n} pairs over {npoly} hulls in {time.time()-t0:.1f}s", flush=True)
eng = pkg.Engine(np.float32); eng.set_device(0); eng.set_sync(False)
bd, _keep = pkg.make_polytopes(pool)
dp, dc, dpairs, dsimp, ddist, dnrm = eng.allocate_indexed_device(bd, n)
eng.upload_pairs_device(pairs, dpairs)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
def step(epa=True):
    eng.compute_minimum_distance_indexed_device(n, dp, dpairs, dsimp, ddist)
    if epa: eng.compute_epa_indexed_device(n, dp, dpairs, dsimp, ddist, dnrm)
for epa in (False, True):
    for _ in range(2): step(epa)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); step(epa); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    print(f"cfg5 n={n} epa={epa} kernel={os.environ.get('OGJK_GJK_KERNEL','auto')}: {ms:.3f} ms {n/ms*1e3:.3e} pairs/s", flush=True)
# ---- device broad phase over the same spheres (grid cell = 2 * max radius, box [0, edge)^3 shifted to +-edge/2)
cell = 5.0
grid = max(1, int(np.ceil(edge / cell)))
sph = spheres.copy(); sph[:, :3] -= edge / 2
d_sph = torch.from_numpy(sph).cuda()
cap = int(n * 1.3) + 1024
d_bp = torch.empty((cap, 2), dtype=torch.int32, device='cuda')
for _ in range(2): total = eng.broadphase_pairs_device(npoly, d_sph, cell, edge / 2, grid, d_bp, cap)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); total = eng.broadphase_pairs_device(npoly, d_sph, cell, edge / 2, grid, d_bp, cap); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"broad phase: {npoly} spheres, grid {grid}^3 -> {total} pairs in {min(ts):.3f} ms ({total/min(ts)*1e3:.3e} pairs/s); host generator found {n}", flush=True)
dist = torch.empty(n, dtype=torch.float32, device='cuda')
import ctypes
ctypes.CDLL('libcudart.so').cudaMemcpy(ctypes.c_void_p(dist.data_ptr()), ctypes.c_void_p(ddist), ctypes.c_size_t(4*n), ctypes.c_int(3))
d = dist.cpu().numpy()
print("colliding fraction:", float((d < 0).mean()), "touching/zero:", float((d == 0).mean()))
eng.free_indexed_device(dp, dc, dpairs, dsimp, ddist, dnrm)

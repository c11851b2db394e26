"""Quick device-resident timing of the GJK / EPA kernels (development helper, not the judged bench)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package
pkg = load_package()

def run(n, nv, spread, dtype, epa=False, reps=5):
    eng = pkg.Engine(dtype); eng.set_device(0); eng.set_sync(False)
    a, b = pkg.workloads.random_pairs(n, nv, spread, seed=12345, dtype=dtype)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    da = torch.from_numpy(a).cuda(); db = torch.from_numpy(b).cuda()
    simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device='cuda')
    dist = torch.zeros(n, dtype=tdt, device='cuda'); nrm = torch.zeros(n, 3, dtype=tdt, device='cuda')
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    def step():
        eng.gjk_uniform_device(n, nv, da, nv, db, simp, dist)
        if epa: eng.epa_uniform_device(n, nv, da, nv, db, simp, dist, nrm)
    for _ in range(3): step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    print(f"n={n} V={nv} S={spread} {np.dtype(dtype).name} epa={epa}: {ms:.3f} ms  {n/ms*1e3:.3e} pairs/s  collide={(dist.cpu().numpy()<=np.finfo(dtype).eps).mean():.3f}", flush=True)

if __name__ == '__main__':
    run(1 << 20, 64, 10.0, np.float32)
    run(1 << 20, 32, 1.0, np.float32, epa=True)
    run(1 << 20, 32, 1.0, np.float32, epa=False)
    run(1 << 18, 64, 10.0, np.float64)
    for nv in (8, 16, 128, 256, 1024):
        run(100000, nv, 10.0, np.float32)

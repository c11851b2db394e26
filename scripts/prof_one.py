"""Runs the device-resident GJK (optionally + EPA) step a few times on one config (target for ncu).
Usage: [OGJK_GJK_KERNEL=...] python scripts/prof_one.py NV SPREAD [N] [REPS] [epa]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package
pkg = load_package()
nv = int(sys.argv[1]); spread = float(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
epa = len(sys.argv) > 5 and sys.argv[5] == 'epa'
dtype = np.float64 if os.environ.get('OGJK_F64') else np.float32
tdt = torch.float64 if dtype == np.float64 else torch.float32
eng = pkg.Engine(dtype); eng.set_device(0); eng.set_sync(False)
a, b = pkg.workloads.random_pairs(n, nv, spread, seed=12345, dtype=dtype)
da = torch.from_numpy(a).cuda(); db = torch.from_numpy(b).cuda()
simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device='cuda')
dist = torch.zeros(n, dtype=tdt, device='cuda'); nrm = torch.zeros(n, 3, dtype=tdt, device='cuda')
eng.set_stream(torch.cuda.current_stream().cuda_stream)
ts = []
for _ in range(reps):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); eng.gjk_uniform_device(n, nv, da, nv, db, simp, dist)
    if epa: eng.epa_uniform_device(n, nv, da, nv, db, simp, dist, nrm)
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"{np.dtype(dtype).name} kernel={os.environ.get('OGJK_GJK_KERNEL','auto')} n={n} V={nv} S={spread} epa={epa}: min {min(ts):.3f} ms {n/min(ts)*1e3:.3e} pairs/s", flush=True)

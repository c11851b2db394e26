"""Experiment helper: GJK on config 2 under a sequence of environment settings (KEY=VALUE ...), results compared with the first."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package, load_oracle
pkg = load_package()
n, nv, spread = 1 << 20, 64, 10.0
dtype = np.float32
eng = pkg.Engine(dtype); eng.set_device(0); eng.set_sync(False)
a, b = pkg.workloads.random_pairs(n, nv, spread, seed=12345, dtype=dtype)
da = torch.from_numpy(a).cuda(); db = torch.from_numpy(b).cuda()
simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device='cuda')
dist = torch.zeros(n, dtype=torch.float32, device='cuda')
eng.set_stream(torch.cuda.current_stream().cuda_stream)
ref = None
for frac in sys.argv[1:]:
    key, val = frac.split('=') if '=' in frac else ('OGJK_CORES', frac)
    os.environ[key] = val
    dist.zero_(); simp.zero_()
    for _ in range(3): eng.gjk_uniform_device(n, nv, da, nv, db, simp, dist)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); eng.gjk_uniform_device(n, nv, da, nv, db, simp, dist); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    d = dist.cpu().numpy(); s = simp.cpu().numpy()
    if ref is None: ref = (d.copy(), s.copy())
    same = np.array_equal(d, ref[0]) and np.array_equal(s, ref[1])
    print(f"{frac}: median {np.median(ts):.3f} ms min {min(ts):.3f} ms  {n/np.median(ts)*1e3:.3e} pairs/s  same_as_first={same}", flush=True)

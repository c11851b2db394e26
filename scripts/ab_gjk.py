"""A/B timing of the GJK kernel families on config 2/3 (development helper). Usage: OGJK_GJK_KERNEL=slots python scripts/ab_gjk.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package, load_oracle
pkg = load_package()
def run(n, nv, spread, check=True):
    dtype = np.float32
    eng = pkg.Engine(dtype); eng.set_device(0); eng.set_sync(False)
    a, b = pkg.workloads.random_pairs(n, nv, spread, seed=12345, dtype=dtype)
    da = torch.from_numpy(a).cuda(); db = torch.from_numpy(b).cuda()
    simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device='cuda')
    dist = torch.zeros(n, dtype=torch.float32, device='cuda')
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    step = lambda: eng.gjk_uniform_device(n, nv, da, nv, db, simp, dist)
    for _ in range(3): step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    ok = ''
    if check:
        m = min(n, 100000)
        orc = load_oracle().Oracle('port', dtype)
        os_, od = orc.gjk(a[:m], b[:m], nthreads=8)
        s = simp.cpu().numpy().view(eng.sdtype)[:m]
        ok = 'dist_eq=%s wit_eq=%s nv_eq=%s' % (np.array_equal(dist.cpu().numpy()[:m], od), np.array_equal(s['witnesses'], os_['witnesses']), np.array_equal(s['nvrtx'], os_['nvrtx']))
    print(f"kernel={os.environ.get('OGJK_GJK_KERNEL','auto')} n={n} V={nv} S={spread}: {ms:.3f} ms {n/ms*1e3:.3e} pairs/s {ok}", flush=True)
if __name__ == '__main__':
    run(1 << 20, 64, 10.0)
    run(1 << 20, 32, 1.0)
    run(1 << 20, 32, 10.0, check=False)
    run(1 << 20, 16, 10.0, check=False)
    run(1 << 20, 8, 10.0, check=False)

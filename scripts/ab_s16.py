"""A/B timing of the fp16 pre-scan slot kernel against the fp32-slot kernel (development helper, GPU).
Device-resident GJK only, median of 10 after 3 warm-ups, CUDA events on the launching stream; every variant's full
output is compared with the first variant's (bit for bit) and the first 100 k pairs with the oracle.
Usage: python scripts/ab_s16.py [NV SPREAD [N]]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package, load_oracle
pkg = load_package()

def run(n, nv, spread, variants):
    dtype = np.float32
    eng = pkg.Engine(dtype); eng.set_device(0); eng.set_sync(False)
    a, b = pkg.workloads.random_pairs(n, nv, spread, seed=12345, dtype=dtype)
    da = torch.from_numpy(a).cuda(); db = torch.from_numpy(b).cuda()
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    m = min(n, 100000)
    os_, od = load_oracle().Oracle('port', dtype).gjk(a[:m], b[:m], nthreads=8)
    first = None
    for name, env in variants:
        for k in ("OGJK_GJK_KERNEL", "OGJK_S16_CFG", "OGJK_S16_IDLE", "OGJK_S16_AGE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device='cuda')
        dist = torch.zeros(n, dtype=torch.float32, device='cuda')
        step = lambda: eng.gjk_uniform_device(n, nv, da, nv, db, simp, dist)
        for _ in range(3): step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        s = simp.cpu().numpy().view(eng.sdtype); d = dist.cpu().numpy()
        ok = np.array_equal(d[:m], od) and np.array_equal(s['witnesses'][:m], os_['witnesses']) and np.array_equal(s['nvrtx'][:m], os_['nvrtx'])
        if first is None:
            first = (s.copy(), d.copy()); same = True
        else:
            same = np.array_equal(d, first[1]) and np.array_equal(s['witnesses'], first[0]['witnesses']) and np.array_equal(s['nvrtx'], first[0]['nvrtx'])
        print(f"{name:22s} n={n} V={nv} S={spread}: {ms:.3f} ms  {n/ms*1e3:.3e} pairs/s  kernel='{eng.last_kernel()}'  oracle_eq={ok} same_as_first={same}", flush=True)

if __name__ == '__main__':
    nv = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    spread = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
    variants = [("fp32 slots (ws)", {"OGJK_GJK_KERNEL": "slotsws32"}), ("fp16 slots K4", {"OGJK_GJK_KERNEL": "slots16"})]
    if nv == 64:
        variants += [(f"fp16 slots K{c}", {"OGJK_GJK_KERNEL": "slots16", "OGJK_S16_CFG": str(c)}) for c in (2, 3, 5)]
    if nv == 32:
        variants.insert(0, ("fp32 slots (self)", {"OGJK_GJK_KERNEL": "slots"}))
    run(n, nv, spread, variants)

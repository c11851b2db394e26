"""BASELINE configs[3]: vertex-count sweep 8..1024, 100k pairs, fp32 and fp64, GJK device-resident (+ roofline fraction)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package
pkg = load_package()
peak = 6545.6
try: peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception: pass
n = 100000
for dtype in (np.float32, np.float64):
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    eng = pkg.Engine(dtype); eng.set_device(0); eng.set_sync(False)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    for nv in (8, 16, 32, 64, 128, 256, 512, 1024):
        a, b = pkg.workloads.random_pairs(n, nv, 10.0, seed=12345, dtype=dtype)
        da = torch.from_numpy(a).cuda(); db = torch.from_numpy(b).cuda()
        simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device='cuda')
        dist = torch.zeros(n, dtype=tdt, device='cuda')
        step = lambda: eng.gjk_uniform_device(n, nv, da, nv, db, simp, dist)
        for _ in range(3): step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        bpp = 2 * nv * 3 * np.dtype(dtype).itemsize + eng.sdtype.itemsize + np.dtype(dtype).itemsize
        gbs = bpp * n / (ms * 1e-3) / 1e9
        print(f"{np.dtype(dtype).name} V={nv:5d} n={n}: {ms:8.3f} ms  {n/ms*1e3:.3e} pairs/s  {gbs:7.1f} GB/s algorithmic = {100*gbs/peak:5.1f}% of HBM peak", flush=True)

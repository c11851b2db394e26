"""The reference's performance sweep (GJK::GPU::testing, examples/gpu/example.cu:258-383; published RTX 4070 numbers in
data/data_32bit_4070) on this build: same two sweeps (1000 pairs x {50..5000} vertices, {50..50000} pairs x 500
vertices), same generator distribution (radius 1..1.5, offsets +-5), same timing definition -- the kernel launch
bracketed by device events for the GPU (examples/common/timer.h), a single-thread loop for the CPU (reference CPU
code, oracle/_ref) -- 10 runs after one warm-up, same CSV header.  SURVEY.md section 8(f) row 4.

    python scripts/sweep_csv.py [out.csv]
"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _pkgpath import load_package, load_oracle

PUBLISHED_4070 = {  # (polytopes, vertices): (GPU_ms, CPU_ms), reference data/data_32bit_4070
    (1000, 50): (0.0917, 0.6559), (1000, 100): (0.1012, 1.2461), (1000, 200): (0.1216, 2.4387),
    (1000, 500): (0.3433, 5.3216), (1000, 1000): (0.4123, 15.2590), (1000, 5000): (6.4943, 105.5536),
    (50, 500): (0.2048, 0.4449), (100, 500): (0.2202, 0.6379), (250, 500): (0.2038, 1.2172),
    (500, 500): (0.2130, 2.8365), (5000, 500): (1.2912, 28.2335), (10000, 500): (2.3429, 56.4362),
    (50000, 500): (10.9461, 294.3638),
}
CASES = [(1000, v) for v in (50, 100, 200, 500, 1000, 5000)] + [(n, 500) for n in (50, 100, 250, 500, 1000, 5000, 10000, 50000)]
RUNS = 10

def main(out, cases=None, runs=RUNS, plot_out=None, check=False):
    """writes `out` in the format of GJK::GPU::testing (example.cu:281, 376) and, when `plot_out` is given, the same rows
    in the column layout of the published data file the reference's plotting/create_plots.py reads (data/data_32bit_4070:
    polytopes,Vertices,GPU_ms,CPU_ms).  check=True compares every run's distances with the CPU result, bit for bit."""
    cases = cases or CASES
    pkg = load_package()
    om = load_oracle()
    orc = om.Oracle("ref" if om.available("ref", np.float32) else "port", np.float32)
    eng = pkg.Engine(np.float32); eng.set_device(0); eng.set_sync(False)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    rows = ["NumPolytopes,NumVertices,CPU_Time_ms,GPU_Time_ms"]
    prow = ["polytopes,Vertices,GPU_ms,CPU_ms"]
    print(f"{'pairs':>7} {'verts':>6} {'CPU ms':>10} {'GPU ms':>9} {'speed-up':>9}   published RTX 4070: GPU ms / CPU ms / speed-up")
    for n, nv in cases:
        cpu_sum = gpu_sum = 0.0
        for run in range(runs):
            a, b = pkg.workloads.random_pairs(n, nv, 10.0, seed=1000 * run + nv + n, dtype=np.float32)
            bd1, _k1 = pkg.make_polytopes(a); bd2, _k2 = pkg.make_polytopes(b)
            h = eng.allocate_and_copy_device_arrays(bd1, bd2)
            d_bd1, d_bd2, d_c1, d_c2, d_simp, d_dist = h
            if run == 0:
                eng.compute_minimum_distance_device(n, d_bd1, d_bd2, d_simp, d_dist); torch.cuda.synchronize()
            t0 = time.perf_counter(); cs, cd = orc.gjk(a, b, nthreads=1); cpu_sum += (time.perf_counter() - t0) * 1e3
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); eng.compute_minimum_distance_device(n, d_bd1, d_bd2, d_simp, d_dist); e1.record()
            torch.cuda.synchronize(); gpu_sum += e0.elapsed_time(e1)
            if check:
                gs, gd = eng.copy_results_from_device(n, d_simp, d_dist)
                assert np.array_equal(gd.view(np.uint32), cd.view(np.uint32)), f"{n} x {nv}: distances differ from the CPU run"
            eng.free_device_arrays(*h)
        cpu, gpu = cpu_sum / runs, gpu_sum / runs
        rows.append(f"{n},{nv},{cpu:.6f},{gpu:.6f}")
        prow.append(f"{n},{nv},{gpu:.4f},{cpu:.4f}")
        pub = PUBLISHED_4070.get((n, nv))
        ptxt = f"{pub[0]:.4f} / {pub[1]:.4f} / {pub[1] / pub[0]:.1f}x" if pub else ""
        print(f"{n:7d} {nv:6d} {cpu:10.4f} {gpu:9.4f} {cpu / gpu:8.1f}x   {ptxt}", flush=True)
    open(out, "w").write("\n".join(rows) + "\n")
    if plot_out:
        open(plot_out, "w").write("\n".join(prow) + "\n")
    print("wrote", out, plot_out or "")
    return rows, prow

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/data_32bit_b200.csv",
         plot_out=sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/data_32bit_b200_plot_format.csv")

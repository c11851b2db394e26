#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched GJK distance + EPA penetration hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3]

One "step" = one pass of the hot path (GJK on every pair, then EPA: penetration/witness/normal for the colliding
pairs and the witness normal for the rest = the reference's computeGJKAndEPA, examples/gpu/example.cu:54-84) over
one batch of seeded synthetic pairs.  Default workload = BASELINE.json configs[1]: 1 Mi pairs of 64-vertex random
polytopes, fp32, offsets +-5.  Multi-GPU: the pair array is sharded, every rank
processes its own batch of the same size (weak scaling), no collective on the data path.

Prints ONE JSON line (rank 0).  Keys follow the driver contract; `roofline` is for the dominant kernel (GJK),
`cpu_baseline` is the reference's CPU path (oracle/_ref) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from _pkgpath import load_oracle, load_package  # noqa: E402

WORKLOADS = {
    # name: (pairs per GPU, verts, spread, description)
    "cfg2": (1 << 20, 64, 10.0, "BASELINE configs[1]: 1Mi random convex polytope pairs, 64 verts, fp32, offsets +-5"),
    "cfg3": (1 << 20, 32, 1.0, "BASELINE configs[2]: 1Mi overlapping pairs, 32 verts, fp32, offsets +-0.5"),
}
METRIC = "gjk_epa_collision_pairs_per_sec"
UNIT = "pairs/s"


def algorithmic_bytes_per_pair(nv: int, itemsize: int, simplex_bytes: int) -> int:
    """B_GJK = 2*V*3*sizeof(T) + sizeof(gkSimplex) + sizeof(T)"""
    return 2 * nv * 3 * itemsize + simplex_bytes + itemsize


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------
def reference_arm(args, workload):
    """The reference's own CPU implementation (oracle/_ref: GJK/cpu/openGJK.c + EPA.c compiled unmodified; the C
    port if _ref is absent) on all host threads, each step a bounded sample of the workload."""
    n_full, nv, spread, desc = workload
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    om = load_oracle()
    pkg = load_package()
    kind = "ref" if om.available("ref", np.float32) else "port"
    orc = om.Oracle(kind, np.float32)
    threads = host_threads()
    sample = min(n_full, 1 << 17)
    a, b = pkg.workloads.random_pairs(sample, nv, spread, seed=12345, dtype=np.float32)

    def step():
        s, d = orc.gjk(a, b, nthreads=threads)
        orc.epa(a, b, s, d, nthreads=threads)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "pairs_per_step": sample, "verts": nv, "stage": "gjk+epa"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads,
                         "kind": "reference" if kind == "ref" else "port",
                         "sample": f"first {sample} pairs of the workload per step, OpenMP over pairs"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def ours(args, workload):
    import torch
    import torch.distributed as dist

    n, nv, spread, desc = workload
    if args.pairs:
        n = args.pairs
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    pkg = load_package()
    dtype = np.float32
    eng = pkg.Engine(dtype)  # raises if the CUDA library is missing: there is no fallback
    eng.set_device(local_rank)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    eng.set_sync(False)

    # this rank's shard of the pair array (own seed => distinct pairs per rank)
    a, b = pkg.workloads.random_pairs(n, nv, spread, seed=12345 + rank, dtype=dtype)
    d_a = torch.from_numpy(a).cuda()
    d_b = torch.from_numpy(b).cuda()
    d_simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8, device="cuda")
    d_dist = torch.zeros(n, dtype=torch.float32, device="cuda")
    d_nrm = torch.zeros(n, 3, dtype=torch.float32, device="cuda")

    def step():
        # GJK on every pair, then EPA (penetration / witnesses / normal for the colliding pairs, witness normal for
        # the rest): one library call = the device part of the reference's computeGJKAndEPA
        eng.gjk_epa_uniform_device(n, nv, d_a, nv, d_b, d_simp, d_dist, d_nrm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    eng.launch_count(reset=True)
    eng.set_timing(True)  # the library brackets its GJK and EPA launches with CUDA events on this stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    launches = eng.launch_count()
    total_ms = e0.elapsed_time(e1)
    gjk_sum, epa_sum, calls = eng.stage_times()
    eng.set_timing(False)
    assert calls == args.steps, (calls, args.steps)
    gjk_ms, epa_ms = gjk_sum / calls, epa_sum / calls

    # ---- end to end through the host-pointer API (computeGJKAndEPA semantics), H2D + D2H inside the timed region
    bd1, _k1 = pkg.make_polytopes(torch.from_numpy(a).pin_memory().numpy())
    bd2, _k2 = pkg.make_polytopes(torch.from_numpy(b).pin_memory().numpy())
    h_simp = torch.zeros(n * eng.sdtype.itemsize, dtype=torch.uint8).pin_memory().numpy().view(eng.sdtype)
    h_dist = torch.zeros(n, dtype=torch.float32).pin_memory().numpy()
    h_nrm = torch.zeros(n, 3, dtype=torch.float32).pin_memory().numpy()
    # host->device bandwidth of this box (pinned, 256 MiB), to put the PCIe-bound end-to-end figure in context
    probe = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    d_probe = torch.empty_like(probe, device="cuda")
    d_probe.copy_(probe, non_blocking=True)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    d_probe.copy_(probe, non_blocking=True)
    p1.record()
    torch.cuda.synchronize()
    h2d_gbs = probe.numel() / (p0.elapsed_time(p1) * 1e-3) / 1e9
    del probe, d_probe
    e2e_steps = max(1, min(args.steps, 5))
    eng.compute_gjk_epa(bd1, bd2, h_simp, h_dist, h_nrm)  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.compute_gjk_epa(bd1, bd2, h_simp, h_dist, h_nrm)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None

    # device time of the slowest rank decides (no data-path collective exists; this MAX is the only reduction)
    total_ms, e2e_ms = pkg.sharding.max_over_ranks([total_ms, e2e_s * 1e3], device="cuda")

    if rank == 0:
        sbytes = eng.sdtype.itemsize
        bpp = algorithmic_bytes_per_pair(nv, 4, sbytes)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = bpp * n / (gjk_ms * 1e-3) / 1e9
        traffic = None  # DRAM bytes per GJK launch from the committed ncu capture of this workload, if any
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            key = f"{args.workload}:{n}"
            if key in t:
                traffic = t[key]["dram_bytes_per_launch"]
        except Exception:  # noqa: BLE001
            pass
        value = world * n * args.steps / (total_ms * 1e-3)
        h2d = 2 * n * nv * 3 * 4  # dense uniform batch: coordinates only, descriptors are read on the host
        d2h = n * (sbytes + 4 + 12)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "pairs_per_gpu": n, "verts": nv, "stage": "gjk+epa",
                       "sharding": f"pairs x{world}, no collective", "l2": "inputs (1.5 GB/step) larger than L2"},
            "roofline": {"bound": "hbm", "kernel": "gjk", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "algorithmic_bytes_per_pair": bpp, "kernel_ms": gjk_ms,
                         "timing": "CUDA events recorded by the library around the GJK launch, mean over the timed steps"},
            "kernels_ms": {"gjk": gjk_ms, "epa": epa_ms},
            "gjk_only_pairs_per_sec": n / (gjk_ms * 1e-3),
            "e2e": {"value": world * n * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "api": "ogjk_f32_compute_gjk_epa (host pointers)",
                    "h2d_gbs_measured": h2d_gbs},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(pkg, nv, spread, a, b)
            # second bound of SURVEY.md section 8(d): non-FMA fp32 lane-ops.  F_GJK = I*(2*V*5 + 150) + 100 per pair with
            # I = the pairs' GJK iteration count, taken from the oracle on a sample of this very batch.
            try:
                om = load_oracle()
                m = min(n, 1 << 15)
                _s, _d, it = om.Oracle("port", np.float32).gjk(a[:m], b[:m], nthreads=host_threads(), want_iters=True)
                mean_it = float(np.mean(it))
                flops = mean_it * (2 * nv * 5 + 150) + 100
                mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
                sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
                fp_peak = sms * 128 * mhz * 1e6 / 1e12
                fp_ach = flops * n / (gjk_ms * 1e-3) / 1e12
                out["roofline_fp32"] = {"bound": "fp32 lane-ops, FMA off", "achieved": fp_ach, "peak": fp_peak,
                                        "unit": "TFLOP/s", "frac": fp_ach / fp_peak, "flops_per_pair": flops,
                                        "mean_gjk_iterations": mean_it,
                                        "note": "looser than the HBM bound, which is therefore the roofline reported above"}
            except Exception as e:  # noqa: BLE001
                out["roofline_fp32"] = {"error": str(e)}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(pkg, nv, spread, a, b):
    """reference CPU path (oracle/_ref) on a bounded sample of the same pairs, all host threads"""
    om = load_oracle()
    kind = "ref" if om.available("ref", np.float32) else "port"
    orc = om.Oracle(kind, np.float32)
    threads = host_threads()
    sample = min(a.shape[0], 1 << 17)
    sa, sb = a[:sample], b[:sample]
    best = None
    best1 = None
    for _ in range(3):
        t0 = time.perf_counter()
        s, d = orc.gjk(sa, sb, nthreads=threads)
        orc.epa(sa, sb, s, d, nthreads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    t0 = time.perf_counter()
    s, d = orc.gjk(sa[: 1 << 15], sb[: 1 << 15], nthreads=1)
    orc.epa(sa[: 1 << 15], sb[: 1 << 15], s, d, nthreads=1)
    best1 = time.perf_counter() - t0
    return {"value": sample / best, "unit": UNIT, "cores": threads,
            "kind": "reference" if kind == "ref" else "port",
            "sample": f"first {sample} pairs of the workload, GJK then EPA, OpenMP over pairs, best of 3",
            "one_thread_value": (1 << 15) / best1}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="override pairs per GPU (development)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    workload = WORKLOADS[args.workload]
    if args.impl == "reference":
        reference_arm(args, workload)
    else:
        ours(args, workload)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched GJK distance + EPA penetration hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg5] [--no-extra]

One "step" = one pass of the hot path (GJK on every pair, then EPA: penetration/witness/normal for the colliding
pairs and the witness normal for the rest = the reference's computeGJKAndEPA, examples/gpu/example.cu:54-84) over
one batch of seeded synthetic pairs.

Workloads (BASELINE.json configs):
  cfg2 (default, the headline)  1 Mi pairs of 64-vertex random polytopes, fp32, offsets +-5.  Multi-GPU: every rank
        processes its own 1 Mi batch (weak scaling), no collective on the data path.
  cfg3  1 Mi overlapping pairs of 32-vertex polytopes (offsets +-0.5): EPA on ~98 % of the pairs.  Weak scaling.
  cfg5  ONE pool of 20 000 32-vertex hulls and ONE list of ~16 M broad-phase candidate pairs (gkCollisionPair), GJK+EPA
        through the indexed device API -- the visualiser's per-frame call (integrate_final_gjk.cu:1028-1036).  Multi-GPU:
        STRONG scaling -- the pair list is cut into contiguous slices (sharding.shard_bounds), the pool is replicated,
        every rank computes its slice, the outputs are gathered (all_gather, outside the timed region) and the gathered
        result is checked against the oracle.
The default line carries cfg2 as the contract keys and, under "extra", short runs of cfg3 and cfg5.

After the timed region the outputs of the batch that was timed are compared, pair by pair and bit for bit, with the
reference's CPU code on all host threads: "parity": {"pairs_checked", "mismatches"}.

Prints ONE JSON line (rank 0).  `roofline` is for the dominant kernel of the workload; `cpu_baseline` is the
reference's CPU path (oracle/_ref) on the same pairs; `ref_gpu_baseline` is the reference's own GPU kernels
(oracle/_ref_gpu, recompiled for sm_100) on the same pairs and the same GPU.
"""
from __future__ import annotations

import argparse
import hashlib
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
    # name: (pairs, verts, spread, description)
    "cfg2": (1 << 20, 64, 10.0, "BASELINE configs[1]: 1Mi random convex polytope pairs, 64 verts, fp32, offsets +-5"),
    "cfg3": (1 << 20, 32, 1.0, "BASELINE configs[2]: 1Mi overlapping pairs, 32 verts, fp32, offsets +-0.5"),
    "cfg5": (16_000_000, 32, 0.0, "BASELINE configs[4]: 20000-hull pool (32 verts), ~16M broad-phase candidate pairs, "
                                  "indexed GJK+EPA, pair list sharded across the GPUs"),
}
CFG5_POOL = 20000
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


def config_of(name: str, world: int) -> dict:
    """the `config` object -- identical for our arm and the reference arm"""
    n, nv, _spread, desc = WORKLOADS[name]
    if name == "cfg5":
        return {"workload": desc, "pool": CFG5_POOL, "verts": nv, "stage": "gjk+epa",
                "sharding": "one pair list, contiguous slices, pool replicated, no collective on the hot loop",
                "l2": "outputs (1.9 GB/step) larger than L2; the 7.7 MB pool is meant to stay in L2"}
    return {"workload": desc, "pairs_per_gpu": n, "verts": nv, "stage": "gjk+epa",
            "sharding": "pairs, one batch per GPU, no collective",
            "l2": "inputs (%.1f GB/step) larger than L2" % (2 * n * nv * 12 / 1e9)}


def kernel_source_hash() -> str:
    """hash of the kernel sources: profiles/traffic.json entries are only quoted for the code they were measured on"""
    h = hashlib.sha1()
    csrc = os.path.join(ROOT, "opengjk-gpu_b200", "csrc")
    for f in ("gjk_slots.cuh", "gjk_core.cuh", "gjk_math.cuh", "gjk_tables.h"):
        h.update(open(os.path.join(csrc, f), "rb").read())
    return h.hexdigest()[:12]


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
def checker():
    om = load_oracle()
    kind = "ref" if om.available("ref", np.float32) else "port"
    return om.Oracle(kind, np.float32), ("reference" if kind == "ref" else "port")


def make_workload(pkg, name, rank=0):
    """-> dict with the host arrays of one rank's batch (dense) or the whole job (cfg5)"""
    n, nv, spread, _ = WORKLOADS[name]
    if name == "cfg5":
        pool, pairs = pkg.workloads.broadphase_pool(CFG5_POOL, nv, n)
        return {"pool": pool, "pairs": pairs, "n": int(pairs.shape[0]), "nv": nv}
    a, b = pkg.workloads.random_pairs(n, nv, spread, seed=12345 + rank, dtype=np.float32)
    return {"a": a, "b": b, "n": n, "nv": nv}


def reference_arm(args):
    """The reference's own CPU implementation (oracle/_ref: GJK/cpu/openGJK.c + EPA.c compiled unmodified; the C
    port if _ref is absent) on all host threads, every step the FULL batch of the workload (what one GPU processes per
    step in our arm), same generator, same seed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = load_package()
    orc, kind = checker()
    threads = host_threads()
    w = make_workload(pkg, args.workload)
    n = w["n"]
    if args.workload == "cfg5" and args.pairs:
        n = min(n, args.pairs)

    def step():
        if args.workload == "cfg5":
            orc.gjk_epa_indexed(w["pool"], w["pairs"][:n], nthreads=threads)
        else:
            s, d = orc.gjk(w["a"], w["b"], nthreads=threads)
            orc.epa(w["a"], w["b"], s, d, nthreads=threads)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.workload == "cfg5" else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_of(args.workload, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"the full batch ({n} pairs) per step, GJK then EPA, OpenMP over pairs"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------------------
class Runner:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", self.local_rank))
        torch.cuda.set_device(self.local_rank)
        self.pkg = load_package()
        self.eng = self.pkg.Engine(np.float32)  # raises if the CUDA library is missing: there is no fallback
        self.eng.set_device(self.local_rank)
        self.stream = torch.cuda.current_stream()
        self.eng.set_stream(self.stream.cuda_stream)
        self.eng.set_sync(False)
        self.sbytes = self.eng.sdtype.itemsize

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, step, steps, warmup):
        """-> (total_ms max over ranks, gjk_ms, epa_ms per step, launches)"""
        torch, eng = self.torch, self.eng
        for _ in range(max(warmup, 3)):
            step()
        self.barrier()
        eng.launch_count(reset=True)
        eng.set_timing(True)  # the library brackets its GJK and EPA launches with CUDA events on this stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        for _ in range(steps):
            step()
        e1.record(self.stream)
        self.barrier()
        launches = eng.launch_count()
        total_ms = e0.elapsed_time(e1)
        gjk_sum, epa_sum, calls = eng.stage_times()
        eng.set_timing(False)
        assert calls == steps, (calls, steps)
        (total_ms,) = self.pkg.sharding.max_over_ranks([total_ms], device="cuda")
        return total_ms, gjk_sum / calls, epa_sum / calls, int(launches)

    def sum_over_ranks(self, values):
        torch, dist = self.torch, self.dist
        t = torch.tensor(list(values), dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t.cpu()]

    @staticmethod
    def mismatches(simp, dist, nrm, es, ed, en):
        """pairs whose outputs differ from the reference's in any BIT (so -0.0 against +0.0 counts, and a NaN only matches
        the same NaN)"""
        def bits(x):
            x = np.ascontiguousarray(x)
            return x.view(np.uint32 if x.dtype.itemsize == 4 else np.uint64)
        bad = bits(dist) != bits(ed)
        bad |= (bits(nrm) != bits(en)).any(axis=1)
        bad |= simp["nvrtx"] != es["nvrtx"]
        bad |= (bits(simp["witnesses"]) != bits(es["witnesses"])).reshape(len(ed), -1).any(axis=1)
        return int(bad.sum())

    # ---- dense workloads (cfg2, cfg3): weak scaling --------------------------------------------------------------
    def dense(self, name, steps, warmup, want_e2e, want_baselines):
        torch, eng, pkg = self.torch, self.eng, self.pkg
        w = make_workload(pkg, name, self.rank)
        n = self.args.pairs or w["n"]
        nv = w["nv"]
        a, b = w["a"][:n], w["b"][:n]
        d_a, d_b = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        d_simp = torch.zeros(n * self.sbytes, dtype=torch.uint8, device="cuda")
        d_dist = torch.zeros(n, dtype=torch.float32, device="cuda")
        d_nrm = torch.zeros(n, 3, dtype=torch.float32, device="cuda")

        def step():
            # GJK on every pair, then EPA (penetration / witnesses / normal for the colliding pairs, witness normal for
            # the rest): one library call = the device part of the reference's computeGJKAndEPA
            eng.gjk_epa_uniform_device(n, nv, d_a, nv, d_b, d_simp, d_dist, d_nrm)

        total_ms, gjk_ms, epa_ms, launches = self.timed(step, steps, warmup)
        res = {"n": n, "nv": nv, "total_ms": total_ms, "gjk_ms": gjk_ms, "epa_ms": epa_ms, "launches": launches,
               "value": self.world * n * steps / (total_ms * 1e-3)}
        # ---- parity of the batch that was timed, every pair, against the reference's CPU code
        orc, kind = checker()
        threads = host_threads()
        s, d = orc.gjk(a, b, nthreads=threads)
        es, ed, en = orc.epa(a, b, s, d, nthreads=threads)
        bad = self.mismatches(d_simp.cpu().numpy().view(eng.sdtype), d_dist.cpu().numpy(), d_nrm.cpu().numpy(), es, ed, en)
        bad_all, n_all = self.sum_over_ranks([bad, n])
        res["parity"] = {"pairs_checked": int(n_all), "mismatches": int(bad_all), "checker": kind,
                         "compared": "distance, contact normal, witnesses, nvrtx of every pair, bit for bit"}
        res["colliding_fraction"] = float((ed <= np.finfo(np.float32).eps).mean())
        if want_e2e:
            res["e2e"] = self.e2e_dense(a, b, n, nv, steps, (es, ed, en))
        if want_baselines and self.world == 1 and not self.args.no_cpu:
            res["cpu_baseline"] = cpu_baseline_dense(orc, kind, a, b)
            res["ref_gpu_baseline"] = ref_gpu_baseline(lambda rg: rg.gjk_epa(a, b, do_epa=True, reps=2), n, (ed,))
            res["mean_gjk_iterations"] = mean_iterations(a, b)
        return res

    def e2e_dense(self, a, b, n, nv, steps, want):
        """the same step through the host-pointer API (computeGJKAndEPA semantics): pinned host buffers, H2D of the
        inputs and D2H of the results inside the timed region"""
        torch, eng, pkg = self.torch, self.eng, self.pkg
        bd1, _k1 = pkg.make_polytopes(torch.from_numpy(a).pin_memory().numpy())
        bd2, _k2 = pkg.make_polytopes(torch.from_numpy(b).pin_memory().numpy())
        h_simp = torch.zeros(n * self.sbytes, dtype=torch.uint8).pin_memory().numpy().view(eng.sdtype)
        h_dist = torch.zeros(n, dtype=torch.float32).pin_memory().numpy()
        h_nrm = torch.zeros(n, 3, dtype=torch.float32).pin_memory().numpy()
        h2d_gbs = self.h2d_probe()
        e2e_steps = max(1, min(steps, 5))
        eng.compute_gjk_epa(bd1, bd2, h_simp, h_dist, h_nrm)  # warm-up
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.compute_gjk_epa(bd1, bd2, h_simp, h_dist, h_nrm)
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        (e2e_ms,) = pkg.sharding.max_over_ranks([e2e_ms], device="cuda")
        es, ed, en = want
        bad = self.mismatches(h_simp, h_dist, h_nrm, es, ed, en)
        (bad_all,) = self.sum_over_ranks([bad])
        return {"value": self.world * n * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": 2 * n * nv * 3 * 4,  # coordinates only: descriptors are read on the host
                "d2h_bytes_per_step": n * (self.sbytes + 4 + 12), "steps": e2e_steps,
                "api": "ogjk_f32_compute_gjk_epa (host pointers)", "h2d_gbs_measured": h2d_gbs,
                "parity_mismatches": int(bad_all)}

    def h2d_probe(self):
        """host->device bandwidth of this box (pinned, 256 MiB), to put the PCIe-bound end-to-end figure in context"""
        torch = self.torch
        probe = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        d_probe = torch.empty_like(probe, device="cuda")
        d_probe.copy_(probe, non_blocking=True)
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        d_probe.copy_(probe, non_blocking=True)
        p1.record()
        torch.cuda.synchronize()
        return probe.numel() / (p0.elapsed_time(p1) * 1e-3) / 1e9

    # ---- cfg5: one pair list, strong scaling ---------------------------------------------------------------------
    def cfg5(self, steps, warmup, want_e2e, want_baselines):
        torch, eng, pkg, dist = self.torch, self.eng, self.pkg, self.dist
        w = make_workload(pkg, "cfg5")  # every rank generates the same seeded job (7 s of host time)
        pool, pairs_all, nv = w["pool"], w["pairs"], w["nv"]
        n_all = min(w["n"], self.args.pairs) if self.args.pairs else w["n"]
        pairs_all = pairs_all[:n_all]
        lo, hi = pkg.sharding.shard_bounds(n_all, self.rank, self.world)
        pairs = np.ascontiguousarray(pairs_all[lo:hi])
        n = hi - lo
        desc, _keep = pkg.make_polytopes(pool)
        dp, dc, dpairs, dsimp, ddist, dnrm = eng.allocate_indexed_device(desc, n)
        eng.upload_pairs_device(pairs, dpairs)

        def step():
            eng.gjk_epa_indexed_device(n, dp, dpairs, dsimp, ddist, dnrm)

        try:
            total_ms, gjk_ms, epa_ms, launches = self.timed(step, steps, warmup)
            res = {"n": n_all, "nv": nv, "pairs_this_rank": n, "total_ms": total_ms, "gjk_ms": gjk_ms, "epa_ms": epa_ms,
                   "launches": launches, "value": n_all * steps / (total_ms * 1e-3)}
            # ---- gather the slices (outside the timed region: north_star gathers only outputs, nothing on the hot loop)
            t_simp = raw_view(torch, dsimp, n * self.sbytes)
            t_dist = raw_view(torch, ddist, n * 4).view(torch.float32)
            t_nrm = raw_view(torch, dnrm, n * 12).view(torch.float32).reshape(n, 3)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.barrier()
            g0.record()
            f_simp = pkg.sharding.gather_slices(t_simp.reshape(n, self.sbytes), n_all, self.rank, self.world)
            f_dist = pkg.sharding.gather_slices(t_dist, n_all, self.rank, self.world)
            f_nrm = pkg.sharding.gather_slices(t_nrm, n_all, self.rank, self.world)
            g1.record()
            self.barrier()
            res["gather_ms"] = g0.elapsed_time(g1) if self.world > 1 else 0.0
            # ---- parity of the GATHERED result: rank r checks slice r of the full arrays it received
            orc, kind = checker()
            es, ed, en = orc.gjk_epa_indexed(pool, pairs, nthreads=host_threads())
            got_s = f_simp[lo:hi].cpu().numpy().reshape(-1).view(eng.sdtype)
            bad = self.mismatches(got_s, f_dist[lo:hi].cpu().numpy(), f_nrm[lo:hi].cpu().numpy(), es, ed, en)
            bad_all, cnt_all = self.sum_over_ranks([bad, n])
            res["parity"] = {"pairs_checked": int(cnt_all), "mismatches": int(bad_all), "checker": kind,
                             "compared": "gathered distance, contact normal, witnesses, nvrtx of every pair, bit for bit"}
            res["colliding_fraction"] = float((ed < 0).mean())
            del f_simp, f_dist, f_nrm
            if want_e2e:
                res["e2e"] = self.e2e_cfg5(desc, pool, pairs, n, n_all, nv, steps, (es, ed, en))
            if want_baselines and self.world == 1 and not self.args.no_cpu:
                m = min(n_all, 1 << 21)
                res["cpu_baseline"] = cpu_baseline_indexed(orc, kind, pool, pairs_all[:m])
                res["ref_gpu_baseline"] = ref_gpu_baseline(
                    lambda rg: rg.gjk_epa_indexed(pool, pairs_all, do_epa=True, reps=1), n_all, (ed,))
        finally:
            eng.free_indexed_device(dp, dc, dpairs, dsimp, ddist, dnrm)
        return res

    def e2e_cfg5(self, desc, pool, pairs, n, n_all, nv, steps, want):
        """host-pointer API (compute_gjk_epa_indexed): pool + this rank's pair slice H2D, results D2H, every step"""
        torch, eng, pkg = self.torch, self.eng, self.pkg
        e2e_steps = max(1, min(steps, 3))
        # result arrays in pinned host memory, allocated once (the pair list and the pool are ordinary numpy arrays)
        h_simp = torch.empty(n * self.sbytes, dtype=torch.uint8, pin_memory=True).numpy().view(eng.sdtype)
        h_dist = torch.empty(n, dtype=torch.float32, pin_memory=True).numpy()
        h_nrm = torch.empty((n, 3), dtype=torch.float32, pin_memory=True).numpy()
        out = (h_simp, h_dist, h_nrm)
        eng.compute_gjk_epa_indexed(desc, pairs, out=out)  # warm-up: device buffers of the library's pool get their size
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            s, d, nr = eng.compute_gjk_epa_indexed(desc, pairs, out=out)
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        (e2e_ms,) = pkg.sharding.max_over_ranks([e2e_ms], device="cuda")
        es, ed, en = want
        bad = self.mismatches(s, d, nr, es, ed, en)
        (bad_all,) = self.sum_over_ranks([bad])
        return {"value": n_all * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(pool.nbytes + CFG5_POOL * 32 + n * 8),
                "d2h_bytes_per_step": n * (self.sbytes + 4 + 12), "steps": e2e_steps,
                "api": "ogjk_f32_compute_gjk_epa_indexed (host pointers; results into pinned host arrays, chunked and overlapped with the kernels)",
                "parity_mismatches": int(bad_all)}


def raw_view(torch, ptr, nbytes):
    class _P:
        def __init__(self):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3}

    return torch.as_tensor(_P(), device="cuda")


def cpu_baseline_dense(orc, kind, a, b):
    """reference CPU path on the same pairs (the full batch), all host threads, best of 3; and one thread as shipped"""
    threads = host_threads()
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        s, d = orc.gjk(a, b, nthreads=threads)
        orc.epa(a, b, s, d, nthreads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    m = min(a.shape[0], 1 << 15)
    t0 = time.perf_counter()
    s, d = orc.gjk(a[:m], b[:m], nthreads=1)
    orc.epa(a[:m], b[:m], s, d, nthreads=1)
    one = time.perf_counter() - t0
    return {"value": a.shape[0] / best, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"the full batch ({a.shape[0]} pairs), GJK then EPA, OpenMP over pairs, best of 3",
            "one_thread_value": m / one}


def cpu_baseline_indexed(orc, kind, pool, pairs):
    threads = host_threads()
    t0 = time.perf_counter()
    orc.gjk_epa_indexed(pool, pairs, nthreads=threads)
    dt = time.perf_counter() - t0
    return {"value": pairs.shape[0] / dt, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"first {pairs.shape[0]} pairs of the list, GJK then EPA, OpenMP over pairs"}


def ref_gpu_baseline(run, n, want):
    """the reference's own GPU kernels (GJK/gpu/openGJK.cu recompiled for sm_100, oracle/_ref_gpu) on the same pairs
    and the same GPU; timing = the reference's definition (cudaEvent pair around the *_device calls)"""
    om = load_oracle()
    if not om.RefGpu.available():
        return {"unavailable": "oracle/_ref_gpu not built"}
    try:
        _s, d, _nr, gjk_ms, epa_ms = run(om.RefGpu())
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)}
    (ed,) = want
    return {"gjk_ms": gjk_ms, "epa_ms": epa_ms, "value": n / ((gjk_ms + epa_ms) * 1e-3), "unit": UNIT,
            "gjk_only_pairs_per_sec": n / (gjk_ms * 1e-3),
            "kind": "reference GPU kernels (compute_minimum_distance_device + compute_epa_device), sm_100 recompile",
            "agrees_with_reference_cpu_within_1e-3": float(np.mean(np.abs(d - ed) <= 1e-3)),
            "bit_identical_to_reference_cpu": float(np.mean(d == ed))}


def mean_iterations(a, b):
    try:
        om = load_oracle()
        m = min(a.shape[0], 1 << 15)
        _s, _d, it = om.Oracle("port", np.float32).gjk(a[:m], b[:m], nthreads=host_threads(), want_iters=True)
        return float(np.mean(it))
    except Exception:  # noqa: BLE001
        return None


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        return {}


def roofline_objects(name, res, peaks, clocks, sms):
    """`roofline` (HBM, SURVEY 8d bytes) for the dominant kernel + the fp32 lane-op bound for GJK beside it"""
    n, nv = res["n"], res["nv"]
    peak = float(peaks.get("hbm_gbs", 6650.0))
    src = "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"
    out = {}
    if name == "cfg5":
        # indexed: the pool is L2-resident, HBM sees pair record + simplex + distance + normal = 132 B per pair
        bpp = 8 + 108 + 4 + 12
        ms = res["gjk_ms"] + res["epa_ms"]
        ach = bpp * res["pairs_this_rank"] / (ms * 1e-3) / 1e9
        out["roofline"] = {"bound": "hbm", "kernel": "gjk+epa (indexed)", "achieved": ach, "peak": peak, "unit": "GB/s",
                           "frac": ach / peak, "traffic": None, "peak_source": src, "algorithmic_bytes_per_pair": bpp,
                           "kernel_ms": ms, "note": "EPA is issue-bound, not bandwidth-bound: see kernels_ms and profiles/"}
        return out
    bpp = algorithmic_bytes_per_pair(nv, 4, 108)
    dominant = "gjk" if res["gjk_ms"] >= res["epa_ms"] else "epa"
    ach = bpp * n / (res["gjk_ms"] * 1e-3) / 1e9
    traffic = None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = t.get(f"{name}:{n}")
        if ent and ent.get("source_hash") == kernel_source_hash():
            traffic = ent["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        pass
    out["roofline"] = {"bound": "hbm", "kernel": "gjk", "dominant_kernel_of_step": dominant, "achieved": ach, "peak": peak,
                       "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": src,
                       "algorithmic_bytes_per_pair": bpp, "kernel_ms": res["gjk_ms"],
                       "timing": "CUDA events recorded by the library around the GJK launch, mean over the timed steps"}
    it = res.get("mean_gjk_iterations")
    if it:
        flops = it * (2 * nv * 5 + 150) + 100
        mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
        fp_peak = sms * 128 * mhz * 1e6 / 1e12
        fp_ach = flops * n / (res["gjk_ms"] * 1e-3) / 1e12
        out["roofline_fp32"] = {"bound": "fp32 lane-ops, FMA off", "achieved": fp_ach, "peak": fp_peak, "unit": "TFLOP/s",
                                "frac": fp_ach / fp_peak, "flops_per_pair": flops, "mean_gjk_iterations": it,
                                "note": "looser than the HBM bound, which is therefore the roofline reported above"}
    return out


def summary(res):
    keep = ("value", "total_ms", "gjk_ms", "epa_ms", "launches", "parity", "colliding_fraction", "gather_ms", "e2e",
            "cpu_baseline", "ref_gpu_baseline", "n")
    return {k: res[k] for k in keep if k in res}


def ours(args):
    R = Runner(args)
    torch = R.torch
    name = args.workload
    sampler = ClockSampler(R.local_rank) if R.rank == 0 else None
    if sampler:
        sampler.start()
    if name == "cfg5":
        res = R.cfg5(args.steps, args.warmup, True, True)
    else:
        res = R.dense(name, args.steps, args.warmup, True, True)
    clocks = sampler.stop() if sampler else None
    extra = {}
    if not args.no_extra and not args.pairs:
        xs = max(3, min(args.steps, 5))
        for other in ("cfg3", "cfg5"):
            if other == name:
                continue
            try:
                r = R.cfg5(xs, 3, True, R.world == 1) if other == "cfg5" else R.dense(other, xs, 3, True, R.world == 1)
                r["ms_per_step"] = r["total_ms"] / xs
                r["steps"] = xs
                r["scaling"] = "strong" if other == "cfg5" else "weak"
                r["config"] = config_of(other, R.world)
                extra[other] = {**summary(r), "ms_per_step": r["ms_per_step"], "steps": xs, "scaling": r["scaling"],
                                "config": r["config"]}
            except Exception as e:  # noqa: BLE001
                extra[other] = {"error": repr(e)}
            torch.cuda.empty_cache()
    if R.rank == 0:
        peaks = load_peaks()
        sms = torch.cuda.get_device_properties(R.local_rank).multi_processor_count
        out = {
            "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": R.world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": res["total_ms"] / args.steps, "higher_is_better": True,
            "scaling": "strong" if name == "cfg5" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(name, R.world),
            **roofline_objects(name, res, peaks, clocks, sms),
            "kernels_ms": {"gjk": res["gjk_ms"], "epa": res["epa_ms"]},
            "gjk_only_pairs_per_sec": (res.get("pairs_this_rank") or res["n"]) * R.world / (res["gjk_ms"] * 1e-3),
            "parity": res["parity"],
            "e2e": res.get("e2e"),
            "gpu_launches": res["launches"],
            "clocks": clocks,
        }
        for k in ("cpu_baseline", "ref_gpu_baseline", "gather_ms", "colliding_fraction"):
            if k in res:
                out[k] = res[k]
        if extra:
            out["extra"] = extra
        print(json.dumps(out), flush=True)
    if R.world > 1:
        R.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="override the number of pairs (development)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / ref_gpu_baseline legs")
    ap.add_argument("--no-extra", action="store_true", help="skip the short cfg3 / cfg5 runs reported under 'extra'")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()

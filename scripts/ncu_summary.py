#!/usr/bin/env python
"""Summarise an .ncu-rep (key metrics + instruction mix + hottest source lines) as text for profiles/."""
import collections, csv, io, subprocess, sys

def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout

def main(rep):
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "smsp__average_warp_latency_per_inst_issued.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    print(f"== {rep}")
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                print(f"{h} [{units[i]}] = {vals[i]}")
    # stall reasons
    st = [(h, float(vals[i])) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and vals[i]]
    print("-- stall reasons (warps per issue-active cycle):")
    for h, v in sorted(st, key=lambda x: -x[1])[:8]:
        print(f"   {h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')}: {v:.3f}")
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
    h2 = src[1]
    ia, isrc, it = h2.index("Instructions Executed"), h2.index("Source"), h2.index("Thread Instructions Executed")
    ops, thr, tot = collections.Counter(), collections.Counter(), 0
    for r in src[2:]:
        try:
            n, t = int(r[ia]), int(r[it])
        except (ValueError, IndexError):
            continue
        parts = r[isrc].split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        op = op.split(".")[0]
        ops[op] += n; thr[op] += t; tot += n
    print(f"-- instruction mix (warp instructions executed, total {tot}):")
    for op, n in ops.most_common(16):
        print(f"   {op:8s} {100*n/tot:5.1f}%  avg active lanes {thr[op]/max(n,1):4.1f}")
    cs = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]))))
    cur, agg, txt, hd = None, collections.Counter(), {}, None
    for r in cs:
        if not r: continue
        if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
        if r[0] == "Line No": hd = r; ja = hd.index("Instructions Executed"); continue
        if hd is None: continue
        try: ln = int(r[0]); n = int(r[ja])
        except (ValueError, IndexError): continue
        agg[(cur, ln)] += n; txt[(cur, ln)] = r[1].strip()[:100]
    print("-- hottest source lines (inclusive of inlined callees):")
    for (f, l), n in agg.most_common(14):
        print(f"   {f}:{l}  {100*n/tot:5.1f}%  {txt[(f,l)]}")

if __name__ == "__main__":
    for rep in sys.argv[1:]:
        main(rep)

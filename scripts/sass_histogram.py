#!/usr/bin/env python
"""SASS opcode histograms of the shipped libopengjk_b200.so (cuobjdump -sass, sm_100a): totals over the whole library, then
one block per default kernel.  No GPU needed.   python scripts/sass_histogram.py > profiles/<tag>_sass_histogram.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "opengjk-gpu_b200", "lib", "libopengjk_b200.so")
# the kernels the default policy launches on the BASELINE configurations (demangled-name prefixes)
DEFAULT = [
    "void ogjk::gjk_slots_ws_kernel<float, 4, 1, true, false>",      # config 2 GJK (64+64 vertices, dense)
    "void ogjk::gjk_slots_kernel<float, true, false>",                # configs 3 and 5 GJK (32+32 vertices: self-service slots)
    "void ogjk::gjk_slots_kernel<float, true, true>",                 # opt-in SoA-4 packed pool (OGJK_POOL_PACK=1)
    "void ogjk::epa_queue_kernel<float, ogjk::UniformSource<float> >",  # config 2 EPA (warp per pair) + overflow pass
    "void ogjk::epa_group_kernel_regs<float, 4, 8, ogjk::EpaWorkLean<float>, 96, 4, ogjk::UniformSource<float> >",  # config 3 EPA
    "void ogjk::epa_group_kernel_regs<float, 4, 8, ogjk::EpaWorkLean<float>, 96, 4, ogjk::IndexedSource<float> >",  # config 5 EPA
    "void ogjk::epa_group_kernel<float, 4, 4, ogjk::EpaWorkTiny<float>, 20, 4, ogjk::UniformSource<float> >",  # <= 16 vertices
    "void ogjk::epa_group_kernel<double, 4, 8, ogjk::EpaWorkSmall<double>, 8, 2, ogjk::UniformSource<double> >",  # fp64
    "void ogjk::gjk_generic_kernel<float, 8, ogjk::DescSource<float> >",
    "void ogjk::gjk_uniform_kernel<float, 8, 8>",
    "void ogjk::gjk_slots16_kernel<8, 8, false, 4>",                   # opt-in fp16 pre-scan kernel
]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    for k, d in zip(kernels, dem):
        names[k] = d
    total = collections.Counter()
    for c in kernels.values():
        total.update(c)
    print("# SASS opcode histograms of the shipped libopengjk_b200.so (cuobjdump -sass, sm_100a; scripts/sass_histogram.py), totals")
    print("# over the whole library first, then one block per default kernel.  UBLKCP = TMA bulk copy, SYNCS = mbarrier,")
    print("# FMUL2/FADD2/FFMA2 = packed fp32, HFMA2/HMUL2/HMNMX2 = packed fp16 (pre-scan kernel), REDUX/CREDUX/MATCH/VOTE = warp")
    print("# collectives.  The default kernels' scans contain no FFMA2 (unfused arithmetic); the FFMA2 of")
    print("# gjk_slots_kernel<float,true,true> are the p*1+q packed adds of the opt-in SoA-4 scan.\n")
    print(f"== whole library: {len(kernels)} kernels, {sum(total.values())} instructions")
    for op, n in total.most_common(60):
        print(f"   {op:<14} {n}")
    for want in DEFAULT:
        for k, c in kernels.items():
            if names[k].startswith(want):
                print(f"\n== {names[k][:160]}")
                print(f"   {sum(c.values())} instructions")
                print("   " + "  ".join(f"{op} {n}" for op, n in c.most_common(45)))
                break
        else:
            print(f"\n== {want}: not in the library", file=sys.stderr)


if __name__ == "__main__":
    main()

#!/usr/bin/env bash
# scanner-warp GJK kernel: parity, A/B against the owner-scans form, ncu captures of it and of the sub-warp EPA kernel
tag="${1:-r2c}"
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_slots.py -m gpu -x -q > $out/${tag}_pytest_slots.txt 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_slots.txt
tail -8 $out/${tag}_pytest_slots.txt
{
  for sc in 0 1; do
    OGJK_WS_SC=$sc timeout 120 python scripts/prof_one.py 64 10 1048576 8
    OGJK_WS_SC=$sc timeout 120 python scripts/prof_one.py 96 10 524288 8
    OGJK_WS_SC=$sc timeout 120 python scripts/prof_one.py 48 10 1048576 8
    OGJK_WS_SC=$sc OGJK_GJK_KERNEL=slotsws timeout 120 python scripts/prof_one.py 32 10 1048576 8
  done
} > $out/${tag}_ab_scanners.txt 2>&1
cat $out/${tag}_ab_scanners.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gjk_slots_ws -s 2 -c 1 -f -o $out/${tag}_gjk_sc_cfg2 \
  python scripts/prof_one.py 64 10 1048576 4 > $out/${tag}_ncu_gjk.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:epa_group -s 1 -c 1 -f -o $out/${tag}_epa_small4_cfg3 \
  python scripts/prof_one.py 32 1 1048576 3 epa > $out/${tag}_ncu_epa.log 2>&1
timeout 400 python bench.py --steps 10 --no-extra > $out/${tag}_bench.json 2> $out/${tag}_bench.err
cat $out/${tag}_bench.json | cut -c1-1500
echo done

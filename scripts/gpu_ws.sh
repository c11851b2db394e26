#!/usr/bin/env bash
# fp32 slot kernels after a change to the scan: parity + timing on configs 2/3 and other vertex counts, optional ncu
tag="${1:-r2s}"; out=gpurun_out; mkdir -p $out
timeout -s KILL 600 python -m pytest tests/test_gpu_slots.py tests/test_gpu_degenerate.py -m gpu -x -q > $out/${tag}_pytest_slots.txt 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_slots.txt; tail -4 $out/${tag}_pytest_slots.txt
for k in auto slotsws slots; do
  if [ $k = auto ]; then unset OGJK_GJK_KERNEL; else export OGJK_GJK_KERNEL=$k; fi
  timeout -s KILL 300 python scripts/ab_gjk.py >> $out/${tag}_ab_gjk.txt 2>&1
done
unset OGJK_GJK_KERNEL
for nv in 48 96 128; do timeout -s KILL 120 python scripts/prof_one.py $nv 10 1048576 6 >> $out/${tag}_ab_gjk.txt 2>&1; done
cat $out/${tag}_ab_gjk.txt
if [ "${2:-}" = "prof" ]; then
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gjk_slots_ws -s 2 -c 1 -f -o $out/${tag}_gjk_cfg2 \
  python scripts/prof_one.py 64 10 1048576 4 > $out/${tag}_ncu_gjk.log 2>&1; tail -2 $out/${tag}_ncu_gjk.log
fi
echo done

#!/usr/bin/env bash
# fp16 pre-scan slot kernel: parity tests, A/B timing against the fp32-slot kernel, one ncu --set full capture.
#   gpurun --timeout 900 -- 'bash scripts/gpu_s16.sh r2f'
tag="${1:-r2f}"
out=gpurun_out
mkdir -p $out
timeout -s KILL 420 python -m pytest tests/test_gpu_slots16.py -m gpu -x -q --durations=8 > $out/${tag}_pytest_s16.txt 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_s16.txt
tail -30 $out/${tag}_pytest_s16.txt
for shape in "64 10" "48 10" "32 1"; do
  timeout -s KILL 200 python scripts/ab_s16.py $shape >> $out/${tag}_ab_s16.txt 2>&1
done
cat $out/${tag}_ab_s16.txt
if [ "${2:-}" != "noprof" ]; then
OGJK_GJK_KERNEL=slots16 timeout -s KILL 240 ncu --set full --clock-control none --import-source on -k regex:gjk_slots16 -s 2 -c 1 -f \
  -o $out/${tag}_gjk_s16_cfg2 python scripts/prof_one.py 64 10 1048576 4 > $out/${tag}_ncu_s16.log 2>&1
tail -3 $out/${tag}_ncu_s16.log
fi
echo done

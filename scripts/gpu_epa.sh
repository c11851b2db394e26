#!/usr/bin/env bash
# EPA kernels: parity tests of every kernel family, A/B timing, optional ncu capture of the default kernel on config 3.
#   gpurun --timeout 900 -- 'bash scripts/gpu_epa.sh r2p [prof]'
tag="${1:-r2p}"
out=gpurun_out
mkdir -p $out
timeout -s KILL 600 python -m pytest tests/test_gpu_epa.py tests/test_gpu_degenerate.py tests/test_gpu_slots.py -m gpu -x -q -k "epa or EPA or fused" > $out/${tag}_pytest_epa.txt 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_epa.txt
tail -6 $out/${tag}_pytest_epa.txt
timeout -s KILL 400 python scripts/ab_epa.py > $out/${tag}_ab_epa.txt 2>&1
cat $out/${tag}_ab_epa.txt
if [ "${2:-}" = "prof" ]; then
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:epa_group -s 1 -c 1 -f -o $out/${tag}_epa_cfg3 \
  python scripts/prof_one.py 32 1 1048576 3 epa > $out/${tag}_ncu_epa.log 2>&1
tail -2 $out/${tag}_ncu_epa.log
fi
echo done

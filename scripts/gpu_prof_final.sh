#!/usr/bin/env bash
# launch list of the bench command + ncu --set full captures of the default GJK (cfg2) and EPA (cfg3, cfg2) kernels.
# gpurun copies at most 64 MiB back: call once per part.   bash scripts/gpu_prof_final.sh <tag> gjk|epa
tag="${1:-r2z}"; part="${2:-gjk}"; out=gpurun_out; mkdir -p $out
if [ "$part" = gjk ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $out/${tag}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gjk_slots_ws -s 2 -c 1 -f -o $out/${tag}_gjk_cfg2 \
  python scripts/prof_one.py 64 10 1048576 4 > $out/${tag}_ncu_gjk.log 2>&1
tail -n 2 $out/${tag}_ncu_gjk.log
else
timeout 300 ncu --set full --clock-control none --import-source on -k regex:epa_group -s 1 -c 1 -f -o $out/${tag}_epa_cfg3 \
  python scripts/prof_one.py 32 1 1048576 3 epa > $out/${tag}_ncu_epa.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:epa_queue -s 1 -c 1 -f -o $out/${tag}_epa_cfg2 \
  python scripts/prof_one.py 64 10 1048576 3 epa > $out/${tag}_ncu_epa2.log 2>&1
tail -n 2 $out/${tag}_ncu_epa.log $out/${tag}_ncu_epa2.log
fi
ls -la $out | tail -8
echo done

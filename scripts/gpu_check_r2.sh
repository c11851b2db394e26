#!/usr/bin/env bash
# Round-2 validation call: GPU parity tests (incl. the full-size ones), smoke, the bench line (both arms), the ncu
# launch list of the bench command and one `--set full` capture each of the GJK slot kernel (cfg2) and the EPA queue
# kernel (cfg3).  Output lands in gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_check_r2.sh r2a'
tag="${1:-r2}"
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_gpu.txt 2>&1
nproc >> $out/${tag}_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > $out/${tag}_pytest.txt 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.txt
tail -25 $out/${tag}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.txt 2>&1
echo "smoke exit $?" >> $out/${tag}_smoke.txt
tail -2 $out/${tag}_smoke.txt
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
cat $out/${tag}_bench.json; tail -5 $out/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>&1
cat $out/${tag}_bench_reference_arm.json
if [ "${2:-}" != "noprof" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $out/${tag}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gjk_slots_ws -s 2 -c 1 -f -o $out/${tag}_gjk_cfg2 \
  python scripts/prof_one.py 64 10 1048576 4 > $out/${tag}_ncu_gjk.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:epa_ -s 2 -c 1 -f -o $out/${tag}_epa_cfg3 \
  python scripts/prof_one.py 32 1 1048576 4 epa > $out/${tag}_ncu_epa.log 2>&1
fi
echo done

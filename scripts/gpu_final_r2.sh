#!/usr/bin/env bash
# Final validation of round 2 (one GPU): all GPU parity tests, smoke, the bench line of both arms, the config 5 bench line,
# the ncu launch list of the bench command and one `--set full` capture of the EPA group kernel on config 3.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_final_r2.sh r2zz'
tag="${1:-r2zz}"
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_gpu.txt 2>&1
nproc >> $out/${tag}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q --durations=10 > $out/${tag}_pytest.txt 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.txt
tail -4 $out/${tag}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.txt 2>&1
echo "smoke exit $?" >> $out/${tag}_smoke.txt
tail -2 $out/${tag}_smoke.txt
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
cat $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>&1
cat $out/${tag}_bench_reference_arm.json
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 > $out/${tag}_bench_cfg5.json 2> $out/${tag}_bench_cfg5.err
cat $out/${tag}_bench_cfg5.json; tail -3 $out/${tag}_bench_cfg5.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $out/${tag}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:epa_group -s 1 -c 1 -f -o $out/${tag}_epa_cfg3 \
  python scripts/prof_one.py 32 1 1048576 3 epa > $out/${tag}_ncu_epa.log 2>&1
tail -n 2 $out/${tag}_ncu_epa.log
echo done

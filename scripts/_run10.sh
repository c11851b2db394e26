set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; cat gpurun_out/bench_cfg2.json
python bench.py --workload cfg3 --steps 10 --no-cpu > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; cat gpurun_out/bench_cfg3.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/b_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:gjk_slots_ws -s 3 -c 1 -f -o gpurun_out/prof_ws64_final python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_f1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:epa_queue -s 3 -c 1 -f -o gpurun_out/prof_epa64_final python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_f2.log 2>&1
OGJK_GJK_KERNEL=uniform python scripts/ab_gjk.py > gpurun_out/ab_uniform.log 2>&1
OGJK_GJK_KERNEL=generic python scripts/ab_gjk.py > gpurun_out/ab_generic.log 2>&1
OGJK_GJK_KERNEL=slots python scripts/ab_gjk.py > gpurun_out/ab_slots.log 2>&1
OGJK_GJK_KERNEL=slotsws python scripts/ab_gjk.py > gpurun_out/ab_slotsws.log 2>&1
python scripts/ab_gjk.py > gpurun_out/ab_auto.log 2>&1
cat gpurun_out/ab_*.log

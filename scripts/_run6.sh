set -x
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -3 gpurun_out/bench_cfg2.err; cat gpurun_out/bench_cfg2.json
python bench.py --workload cfg3 --steps 10 --no-cpu > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; cat gpurun_out/bench_cfg3.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/b_ncu.log 2>&1

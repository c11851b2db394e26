OGJK_GJK_KERNEL=slotsws timeout 120 python scripts/prof_one.py 64 10 > gpurun_out/t_ws.log 2>&1; cat gpurun_out/t_ws.log
timeout 300 python scripts/cfg5_bench.py 4000000 > gpurun_out/cfg5.log 2>&1; tail -3 gpurun_out/cfg5.log
timeout 600 python -m pytest tests/test_gpu_slots.py -x -q 2>&1 | tail -2

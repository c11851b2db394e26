set -x
timeout 600 python -m pytest tests/test_gpu_slots.py tests/test_gpu_device_api.py -x -q > gpurun_out/pytest_idx.log 2>&1
tail -12 gpurun_out/pytest_idx.log
timeout 300 python scripts/cfg5_bench.py 4000000 > gpurun_out/cfg5.log 2>&1; tail -5 gpurun_out/cfg5.log
OGJK_GJK_KERNEL=slotsws timeout 300 python scripts/cfg5_bench.py 4000000 > gpurun_out/cfg5_ws.log 2>&1; tail -4 gpurun_out/cfg5_ws.log
OGJK_GJK_KERNEL=slots timeout 120 python scripts/ab_gjk.py > gpurun_out/ab_slots.log 2>&1; cat gpurun_out/ab_slots.log
OGJK_GJK_KERNEL=slotsws timeout 120 python scripts/prof_one.py 64 10 > gpurun_out/t_ws.log 2>&1; cat gpurun_out/t_ws.log

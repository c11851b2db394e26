set -x
timeout 300 python -m pytest tests/test_gpu_slots.py -x -q > gpurun_out/pytest_slots.log 2>&1
tail -3 gpurun_out/pytest_slots.log
OGJK_GJK_KERNEL=slotsws timeout 120 python scripts/prof_one.py 64 10 > gpurun_out/t_ws.log 2>&1
OGJK_GJK_KERNEL=slots timeout 120 python scripts/ab_gjk.py > gpurun_out/ab_v2.log 2>&1
cat gpurun_out/t_ws.log gpurun_out/ab_v2.log
OGJK_GJK_KERNEL=slotsws timeout 300 ncu --set full --import-source on --clock-control none -k regex:gjk_slots -s 3 -c 1 -f -o gpurun_out/prof_ws64_v3 python scripts/prof_one.py 64 10 > gpurun_out/ncu_ws64.log 2>&1

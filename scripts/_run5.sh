set -x
timeout 300 python -m pytest tests/test_gpu_slots.py -x -q > gpurun_out/pytest_slots.log 2>&1
tail -5 gpurun_out/pytest_slots.log
export OGJK_GJK_KERNEL=slotsws
OGJK_WS_LP=2 timeout 120 python scripts/prof_one.py 64 10 > gpurun_out/t_ws_lp2.log 2>&1
OGJK_WS_LP=1 timeout 120 python scripts/prof_one.py 64 10 > gpurun_out/t_ws_lp1.log 2>&1
cat gpurun_out/t_ws_lp2.log gpurun_out/t_ws_lp1.log
OGJK_WS_LP=2 timeout 300 ncu --set full --import-source on --clock-control none -k regex:gjk_slots -s 3 -c 1 -f -o gpurun_out/prof_ws64_lp2 python scripts/prof_one.py 64 10 > gpurun_out/ncu_ws64lp2.log 2>&1

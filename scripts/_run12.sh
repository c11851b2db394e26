set -x
timeout 600 python -m pytest tests/test_gpu_epa.py tests/test_gpu_device_api.py tests/test_gpu_slots.py -x -q > gpurun_out/pytest_epa.log 2>&1
tail -12 gpurun_out/pytest_epa.log
timeout 200 python scripts/prof_one.py 32 1 1048576 5 epa > gpurun_out/t_cfg3.log 2>&1
timeout 200 python scripts/prof_one.py 64 10 1048576 5 epa > gpurun_out/t_cfg2.log 2>&1
OGJK_EPA_KERNEL=warp timeout 200 python scripts/prof_one.py 32 1 1048576 5 epa > gpurun_out/t_cfg3_warp.log 2>&1
cat gpurun_out/t_cfg3.log gpurun_out/t_cfg2.log gpurun_out/t_cfg3_warp.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:epa_group -s 2 -c 1 -f -o gpurun_out/prof_epag python scripts/prof_one.py 32 1 1048576 3 epa > gpurun_out/ncu_epa.log 2>&1

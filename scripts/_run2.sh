set -x
export OGJK_GJK_KERNEL=slots
python scripts/ab_gjk.py > gpurun_out/ab_slots2.log 2>&1
OGJK_SLOTS_PREFETCH=0 python scripts/prof_one.py 64 10 > gpurun_out/t_slots64_nopf.log 2>&1
python scripts/prof_one.py 64 10 > gpurun_out/t_slots64_pf.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:gjk_slots -s 3 -c 1 -f -o gpurun_out/prof_slots64_v2 python scripts/prof_one.py 64 10 > gpurun_out/ncu_slots64.log 2>&1
unset OGJK_GJK_KERNEL
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/ab_slots2.log gpurun_out/t_slots64_nopf.log gpurun_out/t_slots64_pf.log; tail -5 gpurun_out/pytest_gpu.log

set -x
timeout 300 python -m pytest tests/test_gpu_slots.py -x -q > gpurun_out/pytest_slots.log 2>&1
tail -3 gpurun_out/pytest_slots.log
export OGJK_GJK_KERNEL=slotsws
for pf in 0 64 256 2048; do OGJK_SLOTS_PREFETCH=$pf timeout 120 python scripts/prof_one.py 64 10 > gpurun_out/t_ws_pf$pf.log 2>&1; echo "pf=$pf $(cat gpurun_out/t_ws_pf$pf.log)"; done
OGJK_SLOTS_PREFETCH=256 timeout 300 ncu --set full --import-source on --clock-control none -k regex:gjk_slots -s 3 -c 1 -f -o gpurun_out/prof_ws64_v4 python scripts/prof_one.py 64 10 > gpurun_out/ncu_ws64.log 2>&1

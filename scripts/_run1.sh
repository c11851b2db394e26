set -x
export OGJK_GJK_KERNEL=slots
python scripts/prof_one.py 64 10 > gpurun_out/t_slots64.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:gjk_slots -s 3 -c 1 -f -o gpurun_out/prof_slots64 python scripts/prof_one.py 64 10 > gpurun_out/ncu_slots64.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:gjk_slots -s 3 -c 1 -f -o gpurun_out/prof_slots32 python scripts/prof_one.py 32 10 > gpurun_out/ncu_slots32.log 2>&1
unset OGJK_GJK_KERNEL
python scripts/ab_gjk.py > gpurun_out/ab_auto.log 2>&1
cat gpurun_out/t_slots64.log gpurun_out/ab_auto.log

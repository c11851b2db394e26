OGJK_GJK_KERNEL=slotsws timeout 300 ncu --set full --import-source on --clock-control none -k regex:gjk_slots_ws -s 3 -c 1 -f -o gpurun_out/prof_ws64_feed python scripts/prof_one.py 64 10 > gpurun_out/ncu_ws64.log 2>&1
tail -2 gpurun_out/ncu_ws64.log

export OGJK_EPA_KERNEL=group
OGJK_EPA_GROUP=16 timeout 200 python scripts/prof_one.py 32 1 1048576 5 epa
OGJK_EPA_GROUP=16 timeout 200 python scripts/prof_one.py 64 10 1048576 5 epa
OGJK_EPA_GROUP=16 timeout 300 ncu --set full --import-source on --clock-control none -k regex:epa_group -s 2 -c 1 -f -o gpurun_out/prof_epag16 python scripts/prof_one.py 32 1 1048576 3 epa > gpurun_out/ncu_epa.log 2>&1

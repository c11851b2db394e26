export OGJK_GJK_KERNEL=slotsws
for pf in 0 2048 0 8192; do OGJK_SLOTS_PREFETCH=$pf timeout 120 python scripts/prof_one.py 64 10 > gpurun_out/t_ws_pf$pf.log 2>&1; echo "pf=$pf $(cat gpurun_out/t_ws_pf$pf.log)"; done

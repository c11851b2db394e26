OGJK_GJK_KERNEL=slotsws timeout 120 python scripts/prof_one.py 64 10 2>&1 | tail -1
OGJK_GJK_KERNEL=slots timeout 120 python scripts/ab_gjk.py 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_slots.py -x -q 2>&1 | tail -2

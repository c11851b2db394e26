export OGJK_GJK_KERNEL=slotsws
for c in 32 64 128 256 512; do echo "chunk=$c $(OGJK_WS_CHUNK=$c timeout 120 python scripts/prof_one.py 64 10 2>&1 | tail -1)"; done

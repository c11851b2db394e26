#!/usr/bin/env bash
tag="${1:-r2r}"; out=gpurun_out; mkdir -p $out
timeout -s KILL 400 python -m pytest tests/test_gpu_slots.py -m gpu -x -q -k "indexed" > $out/${tag}_pytest_pack.txt 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_pack.txt; tail -4 $out/${tag}_pytest_pack.txt
timeout -s KILL 400 python scripts/ab_pool_pack.py > $out/${tag}_ab_pool_pack.txt 2>&1; cat $out/${tag}_ab_pool_pack.txt
OGJK_POOL_PACK=1 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gjk_slots_kernel -s 2 -c 1 -f -o $out/${tag}_gjk_pack_cfg5 \
  python scripts/cfg5_bench.py > $out/${tag}_ncu_pack.log 2>&1; tail -3 $out/${tag}_ncu_pack.log
echo done

#!/usr/bin/env bash
tag="${1:-r2v}"; out=gpurun_out; mkdir -p $out
timeout -s KILL 600 python -m pytest tests/test_gpu_slots.py tests/test_gpu_slots16.py tests/test_gpu_midlevel_fast.py tests/test_gpu_fullsize.py -m gpu -x -q -k "indexed or fans_out or cfg5" > $out/${tag}_pytest_idx.txt 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_idx.txt; tail -4 $out/${tag}_pytest_idx.txt
timeout -s KILL 600 python scripts/multi_gpu_host_api.py > $out/${tag}_host_api.txt 2>&1; tail -12 $out/${tag}_host_api.txt
timeout -s KILL 600 python bench.py --workload cfg5 --no-extra --steps 5 --warmup 3 --no-cpu > $out/${tag}_bench_cfg5.json 2> $out/${tag}_bench_cfg5.err; cut -c1-1800 $out/${tag}_bench_cfg5.json; tail -3 $out/${tag}_bench_cfg5.err
echo done

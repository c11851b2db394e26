#!/usr/bin/env bash
# One gpurun call that validates the tree on a B200: GPU parity tests, smoke, the bench line (both arms), fp64 slot
# kernel timings and the ncu launch list of the bench command.  Output lands in gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh r1d'
tag="${1:-check}"
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.txt 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.txt
tail -3 $out/${tag}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.txt 2>&1
echo "smoke exit $?" >> $out/${tag}_smoke.txt
tail -2 $out/${tag}_smoke.txt
timeout 300 python bench.py > $out/${tag}_bench_cfg2.json 2> $out/${tag}_bench_cfg2.err
cat $out/${tag}_bench_cfg2.json
timeout 200 python bench.py --workload cfg3 --no-cpu > $out/${tag}_bench_cfg3.json 2> $out/${tag}_bench_cfg3.err
cat $out/${tag}_bench_cfg3.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>&1
{
  OGJK_F64=1 timeout 120 python scripts/prof_one.py 64 10 524288 5
  OGJK_F64=1 OGJK_GJK_KERNEL=generic timeout 120 python scripts/prof_one.py 64 10 524288 5
  OGJK_F64=1 timeout 120 python scripts/prof_one.py 32 1 1048576 5
  OGJK_F64=1 OGJK_GJK_KERNEL=generic timeout 120 python scripts/prof_one.py 32 1 1048576 5
  OGJK_F64=1 timeout 120 python scripts/prof_one.py 16 10 1048576 5
  OGJK_F64=1 OGJK_GJK_KERNEL=generic timeout 120 python scripts/prof_one.py 16 10 1048576 5
} > $out/${tag}_fp64_slots.txt 2>&1
cat $out/${tag}_fp64_slots.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu > $out/${tag}_ncu_bench.log 2>&1
echo done

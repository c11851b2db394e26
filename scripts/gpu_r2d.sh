#!/usr/bin/env bash
tag="${1:-r2d}"
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_slots.py tests/test_gpu_epa.py tests/test_ref_vis_pinning.py -m gpu -x -q > $out/${tag}_pytest.txt 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.txt
tail -8 $out/${tag}_pytest.txt
timeout 200 python tests/golden/make_vis_golden.py $out/vis_reference_kernels.npz > $out/${tag}_vis_golden.txt 2>&1
tail -3 $out/${tag}_vis_golden.txt
{
  for sc in 0 1; do
    OGJK_WS_SC=$sc timeout 120 python scripts/prof_one.py 64 10 1048576 8
    OGJK_WS_SC=$sc timeout 120 python scripts/prof_one.py 96 10 524288 8
    OGJK_WS_SC=$sc timeout 120 python scripts/prof_one.py 48 10 1048576 8
  done
} > $out/${tag}_ab_scanners.txt 2>&1
cat $out/${tag}_ab_scanners.txt
timeout 400 python scripts/ab_epa.py warp auto > $out/${tag}_ab_epa.txt 2>&1
cat $out/${tag}_ab_epa.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gjk_slots_ws -s 2 -c 1 -f -o $out/${tag}_gjk_sc_cfg2 \
  python scripts/prof_one.py 64 10 1048576 4 > $out/${tag}_ncu_gjk.log 2>&1
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "cfg3 or cfg2_full" > $out/${tag}_pytest_full.txt 2>&1
tail -3 $out/${tag}_pytest_full.txt
echo done

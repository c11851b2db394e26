#!/usr/bin/env bash
# EPA kernel check: parity tests that touch EPA, then the A/B of the kernel families.
tag="${1:-r2b}"
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_epa.py tests/test_gpu_degenerate.py tests/test_gpu_slots.py tests/test_gpu_device_api.py -m gpu -x -q > $out/${tag}_pytest_epa.txt 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_epa.txt
tail -15 $out/${tag}_pytest_epa.txt
timeout 600 python scripts/ab_epa.py > $out/${tag}_ab_epa.txt 2>&1
cat $out/${tag}_ab_epa.txt
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "cfg3 or cfg5 or cfg2" > $out/${tag}_pytest_full.txt 2>&1
tail -5 $out/${tag}_pytest_full.txt
echo done

#!/usr/bin/env bash
# EPA group kernel: parity tests, then the A/B of the run-time variants (scripts/ab_epa_svc.py).
# usage: gpu_epa_svc.sh <tag> [variants ...]
tag="${1:-r2x}"
shift
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_epa.py tests/test_gpu_degenerate.py -m gpu -x -q > $out/${tag}_pytest_epa.txt 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_epa.txt
tail -3 $out/${tag}_pytest_epa.txt
timeout 600 python scripts/ab_epa_svc.py "$@" > $out/${tag}_ab_epa_svc.txt 2>&1
cat $out/${tag}_ab_epa_svc.txt
echo done

#!/usr/bin/env bash
# EPA group kernel: parity tests (under each OGJK_EPA_AREA in $AREAS), then the A/B of the run-time variants.
# usage: [AREAS="a b"] gpu_epa_svc.sh <tag> [variants for scripts/ab_epa_svc.py ...]
tag="${1:-r2x}"
shift
out=gpurun_out
mkdir -p $out
for area in ${AREAS:-default}; do
  OGJK_EPA_AREA=$area timeout 600 python -m pytest tests/test_gpu_epa.py tests/test_gpu_degenerate.py -m gpu -x -q > $out/${tag}_pytest_epa_$area.txt 2>&1
  echo "pytest exit $?" >> $out/${tag}_pytest_epa_$area.txt
  tail -3 $out/${tag}_pytest_epa_$area.txt
done
timeout 600 python scripts/ab_epa_svc.py "$@" > $out/${tag}_ab_epa_svc.txt 2>&1
cat $out/${tag}_ab_epa_svc.txt
echo done

#!/usr/bin/env bash
# EPA group kernel: service batching + two-pass horizon.  Parity tests under both horizon variants, then the A/B.
tag="${1:-r2x}"
out=gpurun_out
mkdir -p $out
for hz in 0 1; do
  OGJK_EPA_HZ=$hz timeout 600 python -m pytest tests/test_gpu_epa.py tests/test_gpu_degenerate.py -m gpu -x -q > $out/${tag}_pytest_epa_hz$hz.txt 2>&1
  echo "pytest exit $?" >> $out/${tag}_pytest_epa_hz$hz.txt
  tail -3 $out/${tag}_pytest_epa_hz$hz.txt
done
OGJK_EPA_HZ=1 timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "cfg3 or cfg5" > $out/${tag}_pytest_full_hz1.txt 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_full_hz1.txt
tail -3 $out/${tag}_pytest_full_hz1.txt
timeout 600 python scripts/ab_epa_svc.py 10:0 24:0 24:1 33:1 24:1:small8 > $out/${tag}_ab_epa_svc.txt 2>&1
cat $out/${tag}_ab_epa_svc.txt
echo done

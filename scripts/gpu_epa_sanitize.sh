#!/usr/bin/env bash
# EPA parity tests (incl. the OGJK_EPA_AREA=small family) + compute-sanitizer memcheck of the EPA group kernel instantiations.
tag="${1:-r2zz3}"
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_epa.py -m gpu -x -q > $out/${tag}_pytest_epa.txt 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_epa.txt
tail -3 $out/${tag}_pytest_epa.txt
timeout 400 compute-sanitizer --tool memcheck python scripts/sanitize_epa_r2.py > $out/${tag}_sanitizer_memcheck.txt 2>&1
tail -12 $out/${tag}_sanitizer_memcheck.txt
echo done

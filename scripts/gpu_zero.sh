#!/usr/bin/env bash
tag="${1:-r2t}"; out=gpurun_out; mkdir -p $out
timeout -s KILL 900 python -m pytest tests/test_gpu_epa.py tests/test_cxx_dropin.py tests/test_gpu_fullsize.py -m gpu -q -k "zero or readme or cfg3 or cfg2 or cfg5 or golden" > $out/${tag}_pytest_zero.txt 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_zero.txt; tail -30 $out/${tag}_pytest_zero.txt

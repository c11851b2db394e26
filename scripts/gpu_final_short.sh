#!/usr/bin/env bash
# Short re-validation after a layout-only change of the EPA work area: EPA + full-size parity tests, smoke, both bench lines.
tag="${1:-r2zz2}"
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_epa.py tests/test_gpu_degenerate.py tests/test_gpu_fullsize.py -m gpu -x -q -k "not cfg4 and not reference_gpu" > $out/${tag}_pytest.txt 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.txt
tail -3 $out/${tag}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.txt 2>&1
echo "smoke exit $?" >> $out/${tag}_smoke.txt
tail -2 $out/${tag}_smoke.txt
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
cut -c1-400 $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 > $out/${tag}_bench_cfg5.json 2> $out/${tag}_bench_cfg5.err
cut -c1-400 $out/${tag}_bench_cfg5.json; tail -3 $out/${tag}_bench_cfg5.err
echo done

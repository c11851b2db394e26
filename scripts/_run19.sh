timeout 600 python -m pytest tests/test_gpu_gjk.py -x -q 2>&1 | tail -3
python bench.py --steps 10 --no-cpu > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; python -c "
import json; j=json.load(open('gpurun_out/bench_e2e.json')); print('value %.3e e2e %.3e h2d %.1f GB/s'%(j['value'], j['e2e']['value'], j['e2e']['h2d_gbs_measured']), j['kernels_ms'])"

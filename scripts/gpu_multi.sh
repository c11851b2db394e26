#!/usr/bin/env bash
# multi-GPU call:  gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_multi.sh N tag'
N="${1:-2}"; tag="${2:-r2m}"
out=gpurun_out
mkdir -p $out
nvidia-smi topo -m > $out/${tag}_topo_${N}.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > $out/${tag}_bench_${N}gpu.json 2> $out/${tag}_bench_${N}gpu.err
cat $out/${tag}_bench_${N}gpu.json | cut -c1-3000; tail -3 $out/${tag}_bench_${N}gpu.err
if [ "${3:-}" = "cfg5" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 10 --warmup 3 --workload cfg5 --no-extra > $out/${tag}_bench_cfg5_${N}gpu.json 2> $out/${tag}_bench_cfg5_${N}gpu.err
cat $out/${tag}_bench_cfg5_${N}gpu.json | cut -c1-3000; tail -3 $out/${tag}_bench_cfg5_${N}gpu.err
fi
timeout 600 python scripts/multi_gpu_host_api.py > $out/${tag}_host_api_${N}gpu.txt 2>&1
tail -8 $out/${tag}_host_api_${N}gpu.txt
timeout 300 python -m pytest tests/test_gpu_midlevel_fast.py -m gpu -q -k fans_out > $out/${tag}_pytest_fanout_${N}gpu.txt 2>&1
tail -3 $out/${tag}_pytest_fanout_${N}gpu.txt
echo done

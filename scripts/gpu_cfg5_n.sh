#!/usr/bin/env bash
# strong-scaling bench line of config 5 on N GPUs:  gpurun --gpus N -- 'bash scripts/gpu_cfg5_n.sh N tag'
N="${1:-2}"; tag="${2:-r2zz}"
out=gpurun_out
mkdir -p $out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 10 --warmup 3 --workload cfg5 --no-extra > $out/${tag}_bench_cfg5_${N}gpu.json 2> $out/${tag}_bench_cfg5_${N}gpu.err
cut -c1-1500 $out/${tag}_bench_cfg5_${N}gpu.json; tail -3 $out/${tag}_bench_cfg5_${N}gpu.err
echo done

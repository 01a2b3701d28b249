#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L | head -8
for n in 8 4; do
  echo "=== bench --gpus $n"
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_r02_n$n.json 2> gpurun_out/bench_r02_n$n.err
  echo "rc=$?"
  tail -c 300 gpurun_out/bench_r02_n$n.err
  python tools/bench_summary.py < gpurun_out/bench_r02_n$n.json 2>&1 | head -5
done

#!/bin/bash
# strong-scaling bench lines of the head on 8 / 4 / 2 GPUs of one box (one 4096-pair sequence split over the ranks)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4 2; do
  echo "=== bench --gpus $n"
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n > gpurun_out/bench_39_n$n.json 2> gpurun_out/bench_39_n$n.err
  echo "rc=$?"
  python tools/bench_summary.py < gpurun_out/bench_39_n$n.json 2>&1 | head -5
done

#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4 2; do
  echo "=== bench --gpus $n"
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_r02_n$n.json 2> gpurun_out/bench_r02_n$n.err
  echo "rc=$?"
  python tools/bench_summary.py < gpurun_out/bench_r02_n$n.json 2>&1 | head -5
done
echo "=== bench N=1"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final_n1.json 2> gpurun_out/bench_r02_final_n1.err
echo "rc=$?"
python tools/bench_summary.py < gpurun_out/bench_r02_final_n1.json 2>&1 | head -12

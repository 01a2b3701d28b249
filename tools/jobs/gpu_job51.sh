#!/bin/bash
# variance of the 8-GPU strong-scaling line (default K), three runs back to back
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for t in 1 2 3; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2961$t bench.py --gpus 8 --no-other > gpurun_out/v8_$t.json 2> gpurun_out/v8_$t.err
  echo "run $t rc=$?: $(python tools/bench_summary.py < gpurun_out/v8_$t.json 2>&1 | head -2 | tr '\n' ' ')"
done

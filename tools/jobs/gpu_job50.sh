#!/bin/bash
# bench.py with the cyclic GC off inside the timed region: 2 ranks, bounded
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 2 --no-other > gpurun_out/gc_2.json 2> gpurun_out/gc_2.err
echo "rc=$? $(python tools/bench_summary.py < gpurun_out/gc_2.json 2>&1 | head -1)"

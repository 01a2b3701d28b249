#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in 0 1; do
  echo "=== PRE3_FIT_CACHE=$v"
  PRE3_FIT_CACHE=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-other > gpurun_out/bench_fc$v.json 2> gpurun_out/bench_fc$v.err
  python tools/bench_summary.py < gpurun_out/bench_fc$v.json 2>&1 | head -2
done
timeout 300 python tools/shard_bench.py 2>&1 | tee gpurun_out/shard_fc.log

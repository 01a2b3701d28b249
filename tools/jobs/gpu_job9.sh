#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "pairs or sequence or ransac or smoke or cfg3" 2>&1 | tail -4
echo "=== bench (fused select)"
timeout 600 python bench.py --steps 20 --warmup 5 --no-other > gpurun_out/bench_r02_i.json 2> gpurun_out/bench_r02_i.err
tail -c 300 gpurun_out/bench_r02_i.err
python tools/bench_summary.py < gpurun_out/bench_r02_i.json 2>&1 | head -4

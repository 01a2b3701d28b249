#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== full GPU suite"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/gpu_suite_h.log
echo "=== bench (fused select)"
timeout 600 python bench.py --steps 20 --warmup 5 --no-other > gpurun_out/bench_r02_h.json 2> gpurun_out/bench_r02_h.err
tail -c 300 gpurun_out/bench_r02_h.err
python tools/bench_summary.py < gpurun_out/bench_r02_h.json 2>&1 | head -4
echo "=== bench (separate select kernels)"
PRE3_NO_FUSED_SELECT=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-other > gpurun_out/bench_r02_h2.json 2> gpurun_out/bench_r02_h2.err
python tools/bench_summary.py < gpurun_out/bench_r02_h2.json 2>&1 | head -4

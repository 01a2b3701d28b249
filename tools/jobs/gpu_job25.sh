#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py tests/test_gpu_callers.py tests/test_gpu_dr_ye.py -x -q 2>&1 | tail -4
for v in 1 0; do
  echo "=== PRE3_RESCORE_V1=$v"
  PRE3_RESCORE_V1=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-other > gpurun_out/bench_rs$v.json 2> gpurun_out/bench_rs$v.err
  python tools/bench_summary.py < gpurun_out/bench_rs$v.json 2>&1 | head -2
  PRE3_RESCORE_V1=$v timeout 300 python tools/match_bench.py 2>&1 | tail -2
done

#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for m in 0 1 2; do
  echo "=== PRE3_SEL_FUSED=$m"
  PRE3_SEL_FUSED=$m timeout 600 python bench.py --steps 20 --warmup 5 --no-other > gpurun_out/bench_sel$m.json 2> gpurun_out/bench_sel$m.err
  python tools/bench_summary.py < gpurun_out/bench_sel$m.json 2>&1 | head -2
done
PRE3_SEL_FUSED=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -x -q -k "ransac or pairs or sequence or cfg3 or cfg1" 2>&1 | tail -3

#!/bin/bash
# fused sequence matcher with fp32 norms: matcher + baseline-shape parity tests, bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/t36.log
python bench.py --no-other > gpurun_out/b36.json 2> gpurun_out/b36.err
python tools/bench_summary.py < gpurun_out/b36.json 2>/dev/null | head -3 | tee -a gpurun_out/t36.log

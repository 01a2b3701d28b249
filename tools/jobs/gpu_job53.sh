#!/bin/bash
# repeatability: the matcher / baseline-shape parity tests five times, the bench check (e2e result = device result) three times
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for i in 1 2 3 4 5; do timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -x -q -m gpu 2>&1 | tail -1; done | tee gpurun_out/rep53.log
for i in 1 2 3; do python bench.py --no-other 2>/dev/null | python tools/bench_summary.py | head -1; done | tee -a gpurun_out/rep53.log

#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== i8 tests"
timeout 600 python -m pytest tests/test_gpu_match_i8.py tests/test_gpu_callers.py -x -q 2>&1 | tail -25 | tee gpurun_out/i8_tests.log
echo "=== match tests"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -x -q -k "siftmatch or sequence or cfg2 or box or gateway" 2>&1 | tail -15 | tee -a gpurun_out/i8_tests.log
echo "=== i8 timing"
timeout 300 python tools/match_i8_bench.py 2>&1 | tee gpurun_out/i8_bench.log

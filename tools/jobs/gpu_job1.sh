#!/bin/bash
# round-2 GPU job 1: eval microbench, CTA-pair matcher bring-up (both K-extension layouts), baseline-shape tests
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(timeout 120 tools/evalbench.bin 20000 454656; timeout 120 tools/evalbench.bin 300 4546560) > gpurun_out/evalbench_a.log 2>&1
cat gpurun_out/evalbench_a.log
for lay in 0 1; do
  echo "=== pair kernel, ext layout $lay"
  PRE3_TC_EXTLAYOUT=$lay timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "siftmatch" 2>&1 | tail -15 | tee gpurun_out/pair_layout$lay.log
done
echo "=== v1 kernel sanity"
PRE3_TC_V1=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "siftmatch" 2>&1 | tail -5
echo "=== match_bench v2 (layout 0), v2 ablations, v1"
PRE3_TC_EXTLAYOUT=0 timeout 200 python tools/match_bench.py 2>&1 | tail -3 | tee gpurun_out/match_bench_v2.log
PRE3_TC_EXP=1 timeout 200 python tools/match_bench.py 2>&1 | tail -3 | tee gpurun_out/match_bench_v2_exp1.log
PRE3_TC_EXP=2 timeout 200 python tools/match_bench.py 2>&1 | tail -3 | tee gpurun_out/match_bench_v2_exp2.log
PRE3_TC_V1=1 timeout 200 python tools/match_bench.py 2>&1 | tail -3 | tee gpurun_out/match_bench_v1.log
echo "=== baseline-shape tests on the v1 matcher"
PRE3_TC_V1=1 timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py -m gpu -q 2>&1 | tail -40 | tee gpurun_out/baseline_tests_a.log

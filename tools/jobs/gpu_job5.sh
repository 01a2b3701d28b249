#!/bin/bash
# round-2 GPU job 5: 512-row CTA-pair matcher: parity, ablations, bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -m gpu -q -x -k "siftmatch or cfg2 or box or sequence or graph" 2>&1 | tail -8 | tee gpurun_out/pair3_parity.log
for e in 0 1 2 3; do
  echo "=== PRE3_TC_EXP=$e"
  PRE3_TC_EXP=$e timeout 200 python tools/match_bench.py 2>&1 | tail -2 | tee gpurun_out/mb5_exp$e.log
done
echo "=== EPI=0 product"
PRE3_TC_EPI=0 timeout 200 python tools/match_bench.py 2>&1 | tail -2 | tee gpurun_out/mb5_epi0.log
echo "=== bench"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_e.json 2> gpurun_out/bench_r02_e.err
tail -c 400 gpurun_out/bench_r02_e.err
python tools/bench_summary.py < gpurun_out/bench_r02_e.json 2>&1 | head -8

#!/bin/bash
# round-2 GPU job 2: full GPU suite on the new kernels, matcher epilogue A/B, bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "=== match_bench: product (EPI=1), EPI=0, epilogue-only EPI=1 / EPI=0, MMA-only"
timeout 200 python tools/match_bench.py 2>&1 | tail -2 | tee gpurun_out/mb2_epi1.log
PRE3_TC_EPI=0 timeout 200 python tools/match_bench.py 2>&1 | tail -2 | tee gpurun_out/mb2_epi0.log
PRE3_TC_EXP=2 timeout 200 python tools/match_bench.py 2>&1 | tail -2 | tee gpurun_out/mb2_exp2_epi1.log
PRE3_TC_EXP=2 PRE3_TC_EPI=0 timeout 200 python tools/match_bench.py 2>&1 | tail -2 | tee gpurun_out/mb2_exp2_epi0.log
PRE3_TC_EXP=1 timeout 200 python tools/match_bench.py 2>&1 | tail -2 | tee gpurun_out/mb2_exp1.log
echo "=== full GPU suite"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/gpu_suite_b.log
echo "=== bench"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err
tail -c 600 gpurun_out/bench_r02_b.err
python tools/bench_summary.py < gpurun_out/bench_r02_b.json 2>&1 | tail -40

#!/bin/bash
# round-2 GPU job 4: matcher ablation 3, full GPU suite, bench with graphs, ncu of the two hot kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
PRE3_TC_EXP=3 timeout 200 python tools/match_bench.py 2>&1 | tail -2 | tee gpurun_out/mb4_exp3.log
echo "=== full GPU suite"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/gpu_suite_d.log
echo "=== bench"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_d.json 2> gpurun_out/bench_r02_d.err
tail -c 400 gpurun_out/bench_r02_d.err
python tools/bench_summary.py < gpurun_out/bench_r02_d.json 2>&1 | tail -40
echo "=== ncu (full set) of k_tc_gemm_pair and k_eval on the profiling driver"
timeout 300 python tools/prof_driver.py cfg2 cfg3 > gpurun_out/prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tc_gemm_pair|k_eval' -c 4 -o gpurun_out/r02_d_prof python tools/prof_driver.py cfg2 cfg3 > gpurun_out/ncu_d.log 2>&1
tail -5 gpurun_out/ncu_d.log

#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== new tests"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "sample_index or ransac" 2>&1 | tail -5
echo "=== bench (plain)"
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02_j.json 2> gpurun_out/bench_r02_j.err || { tail -5 gpurun_out/bench_r02_j.err; exit 1; }
python tools/bench_summary.py < gpurun_out/bench_r02_j.json 2>&1 | head -4
echo "=== ncu launch list"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
wc -l gpurun_out/r02_launches.csv

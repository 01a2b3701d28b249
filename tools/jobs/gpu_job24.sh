#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== build() + smoke()"
timeout 900 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "=== full GPU suite"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/gpu_suite_final.log
echo "=== bench N=1"
timeout 900 python bench.py > gpurun_out/bench_r02_final_n1.json 2> gpurun_out/bench_r02_final_n1.err
echo "rc=$?"; tail -c 300 gpurun_out/bench_r02_final_n1.err
python tools/bench_summary.py < gpurun_out/bench_r02_final_n1.json 2>&1 | head -12
echo "=== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference_arm.json 2> gpurun_out/bench_r02_reference_arm.err
echo "rc=$?"; cut -c1-400 gpurun_out/bench_r02_reference_arm.json
echo "=== ncu launch list of the bench command (our kernels)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py > gpurun_out/ncu_launch.log 2>&1
wc -l gpurun_out/r02_launches.csv

#!/bin/bash
# round-2 GPU job 3: instruction-throughput microbench, full GPU suite, bench with graphs
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 120 tools/pipebench.bin | tee gpurun_out/pipebench.log
echo "=== full GPU suite"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/gpu_suite_c.log
echo "=== bench"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_c.json 2> gpurun_out/bench_r02_c.err
tail -c 600 gpurun_out/bench_r02_c.err
python tools/bench_summary.py < gpurun_out/bench_r02_c.json 2>&1 | tail -40

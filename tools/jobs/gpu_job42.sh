#!/bin/bash
# selection kernels with software-prefetched correspondences: whole GPU suite + bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/t42.log
python bench.py --no-other > gpurun_out/b42.json 2> gpurun_out/b42.err
python tools/bench_summary.py < gpurun_out/b42.json 2>/dev/null | head -3 | tee -a gpurun_out/t42.log
python tools/shard_bench.py 2>&1 | tail -6 | tee -a gpurun_out/t42.log

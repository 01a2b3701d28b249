#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/t47.log
python bench.py > gpurun_out/b47.json 2> gpurun_out/b47.err
python tools/bench_summary.py < gpurun_out/b47.json 2>/dev/null | head -6 | tee -a gpurun_out/t47.log

#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|L3" 
PRE3_DEBUG=1 python tools/e2e_bench.py 4096 2>&1 | tail -12 | tee gpurun_out/e2e45.log
for r in 0 3 5 9; do echo "RAW_EVERY=$r"; PRE3_HOST_F32_RAW_EVERY=$r python tools/e2e_bench.py 4096 2>&1 | tail -2; done | tee -a gpurun_out/e2e45.log

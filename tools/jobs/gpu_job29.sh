#!/bin/bash
# fused sequence matcher: parity against the separate kernels, then per-kernel times at the shard sizes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_sequence" 2>&1 | tail -15
timeout 300 python tools/shard_bench.py 2>&1 | tee gpurun_out/shard_fused.log
PRE3_TC_FUSED=0 timeout 300 python tools/shard_bench.py 2>&1 | tee gpurun_out/shard_unfused.log

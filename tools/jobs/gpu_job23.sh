#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== baseline knobs (NT=64, G=4)"
PRE3_EVP_NT=64 PRE3_TIE_G=4 timeout 300 python tools/shard_bench.py 2>&1 | tee gpurun_out/shard_base.log
echo "=== heuristics"
timeout 300 python tools/shard_bench.py 2>&1 | tee gpurun_out/shard_new.log
echo "=== G sweep at NT heuristics"
for g in 8 16 32; do echo "G=$g"; PRE3_TIE_G=$g timeout 300 python tools/shard_bench.py 2>&1 | tee gpurun_out/shard_g$g.log; done
echo "=== parity"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -x -q -k "ransac or pairs or sequence or cfg3 or cfg1" 2>&1 | tail -3

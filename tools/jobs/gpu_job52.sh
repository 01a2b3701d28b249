#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python tools/shard_bench.py 2>&1 | tail -4 | tee gpurun_out/shard52.log

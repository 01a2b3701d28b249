#!/bin/bash
# pipeline: parity of the chunked form, then the sweep of chunks / stream maps
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipeline or graph_replay or sequence_equals" 2>&1 | tail -8
timeout 600 python tools/pipe_bench.py 4096 512 2>&1 | tee gpurun_out/pipe_sweep.log

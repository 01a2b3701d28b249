#!/bin/bash
# whole GPU suite with the fused sequence matcher as the default path
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/gputests_fused.log

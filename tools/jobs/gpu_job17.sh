#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ekf_update.py tests/test_gpu_ekf.py -x -q 2>&1 | tail -25 | tee gpurun_out/ekf_pred_tests.log

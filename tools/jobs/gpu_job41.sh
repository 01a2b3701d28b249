#!/bin/bash
# ncu --set full with source of k_eval_pairloop and the selection kernels (1024-pair sequence, tools/prof_driver.py cfg3)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_eval_pairloop|k_sel_' -c 4 -f -o gpurun_out/r02_i_eval python tools/prof_driver.py cfg3 > gpurun_out/ncu_41.log 2>&1
tail -3 gpurun_out/ncu_41.log

#!/bin/bash
# ncu --set full with source of k_tc_seq_fused (1024-pair sequence, tools/prof_driver.py cfg3)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python tools/prof_driver.py cfg3 > gpurun_out/prof_plain37.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tc_seq_fused' -c 1 -f -o gpurun_out/r02_h_fused python tools/prof_driver.py cfg3 > gpurun_out/ncu_37.log 2>&1
tail -5 gpurun_out/ncu_37.log
ls -la gpurun_out/*.ncu-rep

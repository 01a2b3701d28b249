#!/bin/bash
# ncu --set full of the fused sequence matcher (one launch, 1024 pairs), with source-level stall samples
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_tc_seq_fused -c 1 -s 1 -f -o gpurun_out/r02_f_prof python tools/prof_driver.py cfg3 > gpurun_out/ncu_f.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_f.log
ls -la gpurun_out/r02_f_prof.ncu-rep

#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python tools/prof_driver.py cfg3 > gpurun_out/prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_eval_pairloop' -c 1 -o gpurun_out/r02_m_prof python tools/prof_driver.py cfg3 > gpurun_out/ncu_m.log 2>&1
tail -3 gpurun_out/ncu_m.log

#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_sel|k_sel_' -c 120 --csv --log-file gpurun_out/sel_launches.csv python bench.py --no-other > gpurun_out/ncu43.log 2>&1
python tools/ncu_summary.py --launches gpurun_out/sel_launches.csv --out gpurun_out/sel43 > /dev/null 2>&1
awk -F, '/^kernel,launches/{f=1} f' gpurun_out/sel43_launches.csv

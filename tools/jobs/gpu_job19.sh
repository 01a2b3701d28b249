#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== bench (plain, same command as the profiled one)"
timeout 600 python bench.py --steps 2 --warmup 3 --no-other > gpurun_out/bench_r02_k.json 2> gpurun_out/bench_r02_k.err || { tail -5 gpurun_out/bench_r02_k.err; exit 1; }
python tools/bench_summary.py < gpurun_out/bench_r02_k.json 2>&1 | head -3
echo "=== ncu --set full on our kernels of the second step"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:^k_ --launch-skip 11 -c 11 -f -o /tmp/r02_p_prof python bench.py --steps 2 --warmup 3 --no-other > gpurun_out/ncu_p.log 2>&1
tail -2 gpurun_out/ncu_p.log | cut -c1-200
ls -la /tmp/r02_p_prof.ncu-rep
ncu -i /tmp/r02_p_prof.ncu-rep --page raw --csv > gpurun_out/r02_p_raw.csv 2>/dev/null
ncu -i /tmp/r02_p_prof.ncu-rep --page details --csv > gpurun_out/r02_p_details.csv 2>/dev/null
sz=$(stat -c %s /tmp/r02_p_prof.ncu-rep)
if [ "$sz" -lt 45000000 ]; then cp /tmp/r02_p_prof.ncu-rep gpurun_out/; fi
ls -la gpurun_out | head -5

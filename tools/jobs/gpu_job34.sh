#!/bin/bash
# fused sequence matcher: sweep of the converters' L2 prefetch distance (PRE3_FZ_PF trips of 14 quads; 0 = off)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for pf in 0 2 4 5 6 9; do
  PRE3_FZ_PF=$pf python bench.py --no-other > gpurun_out/pf_$pf.json 2> gpurun_out/pf_$pf.err
  echo "PF=$pf $(python tools/bench_summary.py < gpurun_out/pf_$pf.json | head -2 | tr '\n' ' ')" | tee -a gpurun_out/pf_sweep.log
done
timeout 600 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3 | tee -a gpurun_out/pf_sweep.log

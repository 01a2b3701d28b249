#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "siftmatch" 2>&1 | tail -3
for e in 0 1 3 4; do
  echo "=== PRE3_TC_EXP=$e"
  PRE3_TC_EXP=$e timeout 200 python tools/match_bench.py 2>&1 | tail -2 | tee gpurun_out/mb6_exp$e.log
done
python - <<'PY'
import importlib, sys
sys.path.insert(0, '.')
pre3 = importlib.import_module("3pre_b200")
c = pre3.Context(0)
print("tmem read GB/s (round-1 microbench):", c.measure_tmem_read(), "-> B/clk/SM", c.measure_tmem_read() * 1e9 / 148 / 1.965e9)
PY

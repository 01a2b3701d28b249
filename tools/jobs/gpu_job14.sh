#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "=== full GPU suite"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/gpu_suite_h.log
echo "=== bench"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_h.json 2> gpurun_out/bench_r02_h.err
tail -c 400 gpurun_out/bench_r02_h.err
python tools/bench_summary.py < gpurun_out/bench_r02_h.json 2>&1 | head -12
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_h.json').read().strip().splitlines()[-1])
print(json.dumps(d["evals"]), json.dumps(d["rooflines"]["eval"]))
PY

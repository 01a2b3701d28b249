#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err
echo "rc=$?"; tail -c 1500 gpurun_out/bench_r02_n2.err
python - <<'PY'
import json
t=open('gpurun_out/bench_r02_n2.json').read().strip()
if t:
    d=json.loads(t.splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","scaling","gpu_launches")})
    print("e2e", d["e2e"]["value"], d["e2e"]["pcie_gb_per_s_whole_job"])
    print("weak", d["config"]["weak_scaling"])
    print("other", json.dumps(d["config"]["other_configs"])[:1200])
    print("kernels", {k: round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
PY

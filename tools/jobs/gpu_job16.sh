#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cov.py -x -q 2>&1 | tail -25 | tee gpurun_out/cov_tests.log
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/cov_bench.log
import importlib, time, numpy as np, torch, sys
sys.path.insert(0, "tests")
pre3 = importlib.import_module("3pre_b200")
from test_oracle_cov_cpu import _scene
from oracle import oracle as orc
ctx = pre3.Context(0)
P, N = 4096, 300
Ya, Yb, R, T = _scene(1, N)
dYa = torch.from_numpy(np.tile(Ya[None], (P, 1, 1))).cuda(); dYb = torch.from_numpy(np.tile(Yb[None], (P, 1, 1))).cuda()
rt = torch.from_numpy(np.tile(np.concatenate([R.T.reshape(-1), T])[None], (P, 1))).cuda()
out = torch.zeros(P, pre3.COV_RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for _ in range(2): ctx.cov_est_ransac_batch_dev(dYa, dYb, rt, 12, out)
ctx.sync()
t0 = time.perf_counter()
for _ in range(3): ctx.cov_est_ransac_batch_dev(dYa, dYb, rt, 12, out)
ctx.sync()
ms = (time.perf_counter() - t0) / 3 * 1e3
t0 = time.perf_counter(); orc.cov_est_ransac_deriv(Ya, Yb, R, T); cpu = (time.perf_counter() - t0) * 1e3
print(f"cov: {P} pairs x {N} support points: {ms:.3f} ms GPU ({P / ms * 1e3:.0f} pairs/s); oracle (1 core, C) {cpu:.3f} ms per pair")
PY

#!/usr/bin/env python
"""pre3_pairs on pinned host buffers (the e2e leg of bench.py) alone:  python tools/e2e_bench.py [P]"""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pre3 = importlib.import_module("3pre_b200")
synth = importlib.import_module("3pre_b200.synth")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
parts = [synth.make_batch_torch(512, 3000 + s0, dev) for s0 in range(0, P, 512)]
host = {}
for k in ("desc1", "desc2", "xyz1", "xyz2"):
    t = torch.cat([p[k] for p in parts])
    host[k] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host[k].copy_(t)
del parts
torch.cuda.synchronize()
hn = {k: v.numpy() for k, v in host.items()}
opts = pre3.make_opts(method=0, k=5, max_iteration=2000, adaptive=True, H=2000, seed=7)
out = (np.zeros(P, pre3.RESULT_DTYPE), np.zeros((P, 512, 2), np.int32), np.zeros((P, 512), np.uint8))
ctx = pre3.Context(0)
for rep in range(4):
    t0 = time.perf_counter()
    ctx.pairs(hn["desc1"], hn["desc2"], hn["xyz1"], hn["xyz2"], opts, out=out)
    dt = time.perf_counter() - t0
    print(f"rep {rep}: {dt*1e3:.1f} ms, {P/dt:.0f} pairs/s, BestFit mean {out[0]['best_fit'].mean():.1f}", flush=True)
ctx.close()

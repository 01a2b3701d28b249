#!/usr/bin/env python
"""eval kernel time against the number of sample sets offered (does the pair loop stop where the reference loop stops?)"""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pre3 = importlib.import_module("3pre_b200")
synth = importlib.import_module("3pre_b200.synth")

def main():
    dev = torch.device("cuda", 0)
    ctx = pre3.Context(0)
    ctx.use_torch_stream()
    P = 4096
    sq = synth.make_sequence_torch(P + 1, 77, dev, K=512, n_corr=300)
    desc, xyz = sq["desc"], sq["xyz"]
    for H in (64, 128, 192, 256, 320, 448, 1000, 2000):
        opts = pre3.make_opts(H=H, seed=9)
        res = torch.zeros(P, 240, dtype=torch.uint8, device=dev)
        m = torch.zeros(P, 512, 2, dtype=torch.int32, device=dev)
        k = torch.zeros(P, 512, dtype=torch.uint8, device=dev)
        for _ in range(2):
            ctx.sequence_dev(desc, xyz, opts, res, m, k)
        ctx.timing_enable(True); ctx.timing_read()
        for _ in range(3):
            ctx.sequence_dev(desc, xyz, opts, res, m, k)
        kt = ctx.timing_read(); ctx.timing_enable(False)
        torch.cuda.synchronize()
        r = np.frombuffer(res.cpu().numpy().tobytes(), dtype=pre3.RESULT_DTYPE)
        print(f"H={H}: eval {kt['eval'][0]/3:.4f} ms select {kt['select'][0]/3:.4f} ms  n_consumed mean {r['n_consumed'].mean():.1f} max {r['n_consumed'].max()} n_iter mean {r['n_iter'].mean():.1f} best_fit mean {r['best_fit'].mean():.1f}", flush=True)
    ctx.close()

if __name__ == "__main__":
    main()

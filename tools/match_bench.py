#!/usr/bin/env python
"""Per-kernel times of the matcher alone at the cfg2 (2048x2048, 256 pairs) and cfg3 (512x512, 4096
pairs) shapes:  python tools/match_bench.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pre3 = importlib.import_module("3pre_b200")
synth = importlib.import_module("3pre_b200.synth")


def main():
    dev = torch.device("cuda", 0)
    ctx = pre3.Context(0)
    ctx.use_torch_stream()
    for name, P, K in (("cfg2", 256, 2048), ("cfg3", 4096, 512)):
        slab = 32 if K == 2048 else 512
        parts = [synth.make_batch_torch(slab, 2000 + s0, dev, K1=K, K2=K, n_corr=(K // 2 if K == 2048 else 300))
                 for s0 in range(0, P, slab)]
        d1 = torch.cat([p["desc1"] for p in parts]).contiguous()
        d2 = torch.cat([p["desc2"] for p in parts]).contiguous()
        del parts
        pairs = torch.zeros(P, K, 2, dtype=torch.int32, device=dev)
        n_out = torch.zeros(P, dtype=torch.int32, device=dev)
        for _ in range(3):
            ctx.siftmatch_batch_dev(d1, d2, pairs, None, n_out, 1.5)
        ctx.timing_enable(True)
        for _ in range(5):
            ctx.siftmatch_batch_dev(d1, d2, pairs, None, n_out, 1.5)
        kt = ctx.timing_read()
        ctx.timing_enable(False)
        flops = 2.0 * K * K * 128 * P
        ms = {k: v[0] / 5 for k, v in kt.items() if v[1]}
        print(name, {k: round(v, 4) for k, v in ms.items()}, "gemm TFLOP/s", round(flops / (ms["match_tc"] * 1e-3) / 1e12, 1),
              "matches/pair", float(n_out.float().mean().item()))
        del d1, d2
        torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()

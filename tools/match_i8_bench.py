#!/usr/bin/env python
"""Per-kernel times of the uint8 matcher (tcgen05 kind::i8, exact) against the exact CUDA-core kernel at the cfg2
(2048x2048, 256 pairs) and cfg3 (512x512, 4096 pairs) shapes:  python tools/match_i8_bench.py"""
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
        d1 = torch.cat([torch.clamp(torch.floor(512.0 * p["desc1"] + 0.5), 0, 255).to(torch.uint8) for p in parts]).contiguous()
        d2 = torch.cat([torch.clamp(torch.floor(512.0 * p["desc2"] + 0.5), 0, 255).to(torch.uint8) for p in parts]).contiguous()
        del parts
        pairs = torch.zeros(P, K, 2, dtype=torch.int32, device=dev)
        ref_pairs = torch.zeros(P, K, 2, dtype=torch.int32, device=dev)
        n_out = torch.zeros(P, dtype=torch.int32, device=dev)
        n_ref = torch.zeros(P, dtype=torch.int32, device=dev)
        for engine, reps in ((0, 5), (1, 1)):
            ctx.set_match_engine(engine)
            out, n = (pairs, n_out) if engine == 0 else (ref_pairs, n_ref)
            for _ in range(2 if engine == 0 else 0):
                ctx.siftmatch_batch_dev(d1, d2, out, None, n, 1.5)
            ctx.timing_enable(True)
            ctx.timing_read()
            for _ in range(reps):
                ctx.siftmatch_batch_dev(d1, d2, out, None, n, 1.5)
            kt = ctx.timing_read()
            ctx.timing_enable(False)
            ms = {k: v[0] / reps for k, v in kt.items() if v[1]}
            ops = 2.0 * K * K * 128 * P
            key = "match_tc" if engine == 0 else "match_exact"
            print(name, "engine", engine, {k: round(v, 4) for k, v in ms.items()}, "TOP/s", round(ops / (ms[key] * 1e-3) / 1e12, 1),
                  "matches/pair", float(n.float().mean().item()))
        ctx.set_match_engine(0)
        torch.cuda.synchronize()
        same = bool((n_out == n_ref).all().item())
        for p in range(0, P, max(1, P // 16)):
            k = int(n_ref[p])
            same = same and bool((pairs[p, :k] == ref_pairs[p, :k]).all().item())
        print(name, "tensor-core rows == exact rows:", same)
        del d1, d2
        torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()

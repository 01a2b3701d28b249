#!/usr/bin/env python
"""Per-kernel times of the sequence step at the pair counts one GPU holds when a 4096-pair sequence is split over
1 / 2 / 4 / 8 GPUs (what bounds strong scaling):  python tools/shard_bench.py"""
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
    torch.cuda.set_stream(torch.cuda.Stream())
    ctx = pre3.Context(0)
    ctx.use_torch_stream()
    ctx.set_graphs(True)
    for P in (512, 1024, 2048, 4096):
        sq = synth.make_sequence_torch(P + 1, 77, dev, K=512, n_corr=300)
        desc, xyz = sq["desc"], sq["xyz"]
        opts = pre3.make_opts(H=2000, seed=9)
        res = torch.zeros(P, 240, dtype=torch.uint8, device=dev)
        m = torch.zeros(P, 512, 2, dtype=torch.int32, device=dev)
        k = torch.zeros(P, 512, dtype=torch.uint8, device=dev)
        for _ in range(5):
            ctx.sequence_dev(desc, xyz, opts, res, m, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ctx.sequence_dev(desc, xyz, opts, res, m, k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        ctx.timing_enable(True)
        ctx.timing_read()
        for _ in range(3):
            ctx.sequence_dev(desc, xyz, opts, res, m, k)
        kt = ctx.timing_read()
        ctx.timing_enable(False)
        print(f"P={P} step {ms:.4f} ms ({P / ms * 1e3:.0f} pairs/s)", {kk: round(v[0] / 3, 4) for kk, v in kt.items()}, flush=True)
        del desc, xyz, sq
    ctx.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Step time of the sequence path against the pipeline settings (pre3_set_pipeline: chunks per call; PRE3_PIPE_MAP:
stage -> stream; PRE3_PIPE_PRIO: stream priorities) at the pair counts one GPU holds when the 4096-pair sequence is
split over 1 / 8 GPUs:  python tools/pipe_bench.py [P ...]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pre3 = importlib.import_module("3pre_b200")
synth = importlib.import_module("3pre_b200.synth")


def time_step(ctx, desc, xyz, opts, res, m, k, iters=20):
    for _ in range(5):
        ctx.sequence_dev(desc, xyz, opts, res, m, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ctx.sequence_dev(desc, xyz, opts, res, m, k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_stream(torch.cuda.Stream())
    Ps = [int(a) for a in sys.argv[1:]] or [4096, 512]
    for P in Ps:
        sq = synth.make_sequence_torch(P + 1, 77, dev, K=512, n_corr=300)
        desc, xyz = sq["desc"], sq["xyz"]
        opts = pre3.make_opts(H=2000, seed=9)
        res = torch.zeros(P, 240, dtype=torch.uint8, device=dev)
        m = torch.zeros(P, 512, 2, dtype=torch.int32, device=dev)
        k = torch.zeros(P, 512, dtype=torch.uint8, device=dev)
        ref = None
        for pmap, prio in (("0123", 1), ("0123", 0), ("0122", 1), ("0112", 1), ("0011", 1), ("0101", 1)):
            os.environ["PRE3_PIPE_MAP"] = pmap
            os.environ["PRE3_PIPE_PRIO"] = str(prio)
            ctx = pre3.Context(0)
            ctx.use_torch_stream()
            ctx.set_graphs(True)
            for chunks in ((0, 2, 3, 4, 6, 8, 12, 16, 32) if pmap == "0123" and prio == 1 else (2, 4, 8, 16)):
                if chunks and P // chunks < 128:
                    continue
                ctx.set_pipeline(chunks)
                ms = time_step(ctx, desc, xyz, opts, res, m, k)
                torch.cuda.synchronize()
                sig = (res.cpu().numpy().tobytes(), m.cpu().numpy().tobytes())
                if ref is None:
                    ref = sig
                same = sig[0] == ref[0]
                print(f"P={P} map={pmap} prio={prio} chunks={chunks:2d}: step {ms:.4f} ms ({P / ms * 1e3:.0f} pairs/s)"
                      f" records_equal={same}", flush=True)
            ctx.close()
        del desc, xyz, sq


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Short single-GPU pass over every kernel family, for ncu (launch list / --set full):

    python tools/prof_driver.py [cfg3] [cfg2] [cfg4] [cfg5] [dr_ye] [frames] [ekf_update]      (default: all)

Sizes are cut down so that ~40 replays per launch stay cheap; shapes per unit are the
BASELINE.json ones (512x512 / 2048x2048 descriptors, N=300 / 20000 correspondences, n=1213)."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pre3 = importlib.import_module("3pre_b200")
synth = importlib.import_module("3pre_b200.synth")
se = importlib.import_module("3pre_b200.synth_ekf")
pd = importlib.import_module("3pre_b200.dist")


def main():
    which = sys.argv[1:] or ["cfg3", "cfg2", "cfg4", "cfg5", "dr_ye", "frames", "ekf_update"]
    dev = torch.device("cuda", 0)
    ctx = pre3.Context(0)
    ctx.use_torch_stream()
    if "cfg3" in which:
        P = 1024
        d = synth.make_sequence_torch(P + 1, 3000, dev)
        opts = pre3.make_opts(method=0, k=5, max_iteration=2000, adaptive=True, H=2000, seed=7)
        res = torch.zeros(P, 240, dtype=torch.uint8, device=dev)
        for _ in range(2):
            ctx.sequence_dev(d["desc"], d["xyz"], opts, res)   # the headline path of bench.py
        ctx.sync()
        del d
    if "cfg2" in which:
        P, K = 64, 2048
        d = synth.make_batch_torch(P, 2000, dev, K1=K, K2=K, n_corr=K // 2)
        pairs = torch.zeros(P, K, 2, dtype=torch.int32, device=dev)
        n_out = torch.zeros(P, dtype=torch.int32, device=dev)
        for _ in range(2):
            ctx.siftmatch_batch_dev(d["desc1"], d["desc2"], pairs, None, n_out, 1.5)
        ctx.sync()
        del d
    if "cfg4" in which:
        Fr = 32
        b = se.make_ekf_frames(Fr, 4000, device=dev, n_id=200)
        b["cam"] = dict(se.CAM)
        li = torch.zeros(Fr, b["F"], dtype=torch.uint8, device=dev)
        res = torch.zeros(Fr, 32, dtype=torch.uint8, device=dev)
        for adaptive in (False, True):
            ctx.ransac_hypotheses_batch_dev(b, pre3.make_ekf_opts(H=256, adaptive=adaptive, seed=11), li, res)
        ctx.sync()
        del b
    if "cfg5" in which:
        N, H = 20000, 131072
        c = synth.make_correspondences(5000, N=N, outlier_ratio=0.6)
        Ya, Yb = torch.from_numpy(c.Ya).to(dev), torch.from_numpy(c.Yb).to(dev)
        opts = pre3.make_opts(method=0, k=5, max_iteration=H + 1, adaptive=False, H=H, seed=5)
        for mode in ("first", "reference"):
            rec, _ = pd.ransac_hypothesis_split(ctx, Ya, Yb, opts, mode=mode, want_mask=False)
        print("cfg5", int(rec["best_fit"]), int(rec["best_sample"]))
    if "dr_ye" in which:   # the code_from_dr_ye variant on the cfg3 shape
        L = importlib.import_module("3pre_b200._lib")
        P = 512
        d = synth.make_sequence_torch(P + 1, 8100, dev)
        opts = pre3.make_opts(method=L.METHOD_DR_YE, k=4, max_iteration=700, adaptive=False, H=700, seed=7)
        res = torch.zeros(P, 240, dtype=torch.uint8, device=dev)
        for _ in range(2):
            ctx.sequence_dev(d["desc"], d["xyz"], opts, res)
        ctx.sync()
        del d
    if "frames" in which:  # frames -> filtered maps / fused per-feature lookup + compaction
        F, K = 512, 512
        g = torch.Generator(device=dev).manual_seed(1)
        sr = torch.rand(F, 176, 720, generator=g, device=dev, dtype=torch.float64) + 1.0
        fr = torch.rand(F, K, 4, generator=g, device=dev, dtype=torch.float64) * 140
        desc = torch.rand(F, K, 128, generator=g, device=dev, dtype=torch.float64)
        o = pre3.make_frame_opts()
        x, y, z = (torch.empty(F, 176, 144, dtype=torch.float64, device=dev) for _ in range(3))
        mc = torch.empty(F, dtype=torch.float64, device=dev)
        xyz = torch.empty(F, K, 3, dtype=torch.float64, device=dev)
        nk = torch.empty(F, dtype=torch.int32, device=dev)
        dout = torch.empty_like(desc)
        for _ in range(2):
            ctx.read_xyz_sr4000_batch_dev(sr, o, x, y, z, mc)
            ctx.features_xyz_batch_dev(sr, o, fr, xyz=xyz, n_keep=nk, desc_in=desc, desc_out=dout)
        ctx.sync()
    if "ekf_update" in which:  # update.m on cfg4-shaped frames (n = 1213, m ~ 320)
        Fr = 16
        b = se.make_ekf_frames(Fr, 4000, device=dev, n_id=200, outlier_ratio=0.2)
        sel = (~b["outlier"]).to(torch.uint8)
        xo, Po = torch.empty_like(b["x"]), torch.empty_like(b["P"])
        torch.cuda.synchronize()
        ctx.ekf_update_batch_dev(b, sel, xo, Po)
        ctx.sync()
    ctx.sync()
    print("launches", ctx.launch_count())
    ctx.close()


if __name__ == "__main__":
    main()

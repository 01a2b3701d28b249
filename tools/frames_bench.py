"""Timing of the frame-step kernels alone: python tools/frames_bench.py"""
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

pre3 = importlib.import_module("3pre_b200")
ctx = pre3.Context(0)
ctx.use_torch_stream()
dev = torch.device("cuda", 0)
print(json.dumps(bench.bench_frames(ctx, pre3, dev), indent=1))
# breakdown of the fused path
F, K = 2048, 512
g = torch.Generator(device=dev).manual_seed(1)
sr = torch.rand(F, 176, 720, generator=g, device=dev, dtype=torch.float64) + 1.0
fr = torch.rand(F, K, 4, generator=g, device=dev, dtype=torch.float64) * 140
desc = torch.rand(F, K, 128, generator=g, device=dev, dtype=torch.float64)
o = pre3.make_frame_opts()
xyz = torch.empty(F, K, 3, dtype=torch.float64, device=dev)
nk = torch.empty(F, dtype=torch.int32, device=dev)
dout = torch.empty_like(desc)
mc = torch.empty(F, dtype=torch.float64, device=dev)
print("conf max only      ms", bench._time_steps(lambda: ctx.read_xyz_sr4000_batch_dev(sr, o, None, None, None, mc), 5, 2))
print("lookup, no desc    ms", bench._time_steps(lambda: ctx.features_xyz_batch_dev(sr, o, fr, xyz=xyz, n_keep=nk), 5, 2))
print("lookup + desc      ms", bench._time_steps(
    lambda: ctx.features_xyz_batch_dev(sr, o, fr, xyz=xyz, n_keep=nk, desc_in=desc, desc_out=dout), 5, 2))

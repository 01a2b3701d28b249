"""Timing of BASELINE config 4 alone: python tools/cfg4_bench.py"""
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
o = bench.bench_cfg4(ctx, pre3, torch.device("cuda", 0), 0)
for m in ("adaptive", "fixed_H"):
    print(m, round(o[m]["frames_per_s"]), {k: round(v["ms_per_step"], 3) for k, v in o[m]["kernels"].items()})

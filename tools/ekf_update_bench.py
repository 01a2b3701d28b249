"""Timing of the EKF partial update kernels: python tools/ekf_update_bench.py [Fr]"""
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
print(json.dumps(bench.bench_ekf_update(ctx, pre3, torch.device("cuda", 0),
                                        Fr=int(sys.argv[1]) if len(sys.argv) > 1 else 64), indent=1))

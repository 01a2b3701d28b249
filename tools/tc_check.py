"""GPU smoke of the tensor-core matcher alone (forced engine), small to full-size shapes."""
import importlib, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pre3 = importlib.import_module("3pre_b200")
synth = importlib.import_module("3pre_b200.synth")
from oracle import oracle as orc

ctx = pre3.Context(0)
ctx.set_match_engine(2)
bad = 0
for seed, (K1, K2) in enumerate([(128, 128), (256, 384), (512, 512), (300, 517), (1, 40), (129, 1), (700, 2048)]):
    fp = synth.make_frame_pair(40 + seed, K1=K1, K2=K2, n_corr=min(K1, K2) // 2)
    for dt in (np.float64, np.float32):
        d1, d2 = fp.desc1.astype(dt), fp.desc2.astype(dt)
        pairs, score = ctx.siftmatch(d1, d2, 1.5)
        op, os_ = orc.siftmatch(d1, d2, 1.5)
        ok = np.array_equal(pairs, op) and np.array_equal(score, os_)
        print(K1, K2, dt.__name__, "n", len(pairs), len(op), "OK" if ok else "MISMATCH", flush=True)
        bad += not ok
print("bad", bad)
sys.exit(1 if bad else 0)

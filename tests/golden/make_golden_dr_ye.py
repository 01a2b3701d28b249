"""Golden vectors for the code_from_dr_ye variant.

The reference ships no vectors for this path and MATLAB / Octave are absent ("parity unpinned"), so the
expected outputs come from the INDEPENDENT numpy / LAPACK restatement (oracle/ref_numpy.py:
vodometry_dr_ye, dr_ye_sampler -- written line by line from M/code_from_dr_ye/ransac_dr_ye.m and
vodometry_dr_ye.m), not from the C oracle the GPU tests compare against bit for bit.
Run from the repo root:  python tests/golden/make_golden_dr_ye.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_numpy as rn  # noqa: E402

synth = importlib.import_module("3pre_b200.synth")

out = {}
cases = [("a", 300, 0.30, 700), ("b", 120, 0.50, 700), ("c", 40, 0.20, 300), ("d", 9, 0.0, 700)]
for name, N, rho, H in cases:
    c = synth.make_correspondences(7000 + len(out), N=N, outlier_ratio=rho)
    draws = synth.make_draws(7100 + len(out), H, N)
    g = rn.vodometry_dr_ye(c.Ya, c.Yb, draws, 700)
    out[f"{name}_Ya"], out[f"{name}_Yb"], out[f"{name}_draws"] = c.Ya, c.Yb, draws
    out[f"{name}_counts"] = g["counts"].astype(np.int32)
    out[f"{name}_mask"] = g["mask"]
    out[f"{name}_R"], out[f"{name}_T"] = g["R"], g["T"]
    out[f"{name}_scalars"] = np.array([g["status"], g["op_num"], g["best_sample"], g["n_loops"],
                                       g["n_iteration_ransac"], g["state"]], np.int64)
    out[f"{name}_stats"] = np.array([g["thr"], g["error_mean"], g["error_std"]])

# the sampler on a recorded uniform stream, with matches that share features (k2 repeats, k1 == k2 values)
rng = np.random.Generator(np.random.PCG64(7200))
match = np.stack([np.sort(rng.choice(60, 30, replace=False)), rng.integers(0, 25, 30)], 1).astype(np.int32)
stream = rng.random(4000)
pos = [0]


def rand():
    v = stream[pos[0]]
    pos[0] += 1
    return v


sets = np.array([rn.dr_ye_sampler(match, rand) for _ in range(200)], np.int32)
out["sampler_match"], out["sampler_stream"], out["sampler_sets"] = match, stream[: pos[0]], sets
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "dr_ye.npz"), **out)
print("wrote", len(out), "arrays;", pos[0], "uniforms consumed")

"""Golden vectors for the frame -> per-feature 3-D point step.

The reference ships no vectors for it and MATLAB (fspecial / imfilter) is absent ("parity unpinned"): the expected
outputs come from the INDEPENDENT scipy.ndimage restatement in oracle/ref_numpy.py (read_xyz_sr4000, sift_extract_xyz),
written from M/read_xyz_sr4000.m, M/inittialize_depth_my_version.m and M/SIFT_extract_save.m.  The frames are stored as
float32-exact values to keep the fixture small.  Run from the repo root:  python tests/golden/make_golden_frames.py
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
# ONE base frame of 721 rows (amplitude block zeroed: it is not read by this step); the 720- and 576-row cases are
# its leading rows
sr, fr = synth.make_sr_frames(9721, 1, 160, rows=721)
sr = sr.astype(np.float32).astype(np.float64)
sr[0, :, 432:576] = 0.0
fr = fr.astype(np.float32).astype(np.float64)
out["sr"] = sr[0].astype(np.float32)
out["frames"] = fr[0].astype(np.float32)
for name, rows in (("a", 720), ("b", 576), ("c", 721)):
    m = sr[0][:, :rows].T
    xyz, remain = rn.sift_extract_xyz(m, fr[0].T)
    x, y, z, _ = rn.read_xyz_sr4000(m)
    out[f"{name}_xyz"] = xyz
    out[f"{name}_remain"] = remain.astype(np.int32)
    out[f"{name}_zrow"] = z[70].copy()      # one row of each filtered map
    out[f"{name}_xcol"] = x[:, 0].copy()    # and one border column (zero padding)
    xr, yr, zr, _ = rn.read_xyz_sr4000(m, 1.0, "replicate")
    out[f"{name}_ycol_dr_ye"] = yr[:, 175].copy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "frames.npz"), **out)
print("wrote", len(out), "arrays")

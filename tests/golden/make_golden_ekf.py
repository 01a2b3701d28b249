"""Generates tests/golden/ekf_frames.npz: small synthetic EKF frames (inputs) together with the
outputs of the independent numpy/LAPACK restatement of ransac_hypotheses /
compute_hypothesis_support_fast (oracle/ref_numpy_ekf.py, written from the reference .m files with
MATLAB-shaped dense algebra).

The reference ships no vectors for this path and MATLAB/Octave are absent, so these are NOT
reference outputs ("parity unpinned"): they pin the C oracle and the CUDA path against
the second, independently written restatement, and guard both against drift.

    python tests/golden/make_golden_ekf.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_numpy_ekf as rn  # noqa: E402

se = importlib.import_module("3pre_b200.synth_ekf")

CASES = {  # name -> generator arguments, selections
    "id_only": dict(n_id=24, n_euc=0, H=60),
    "mixed": dict(n_id=16, n_euc=8, interleave=True, H=60),
    "few_ic": dict(n_id=10, n_euc=0, drop_ic=0.75, H=20),
    "missing_z": dict(n_id=20, n_euc=6, drop_z=0.3, H=40),
}


def main():
    out = {}
    for ci, (name, kw) in enumerate(CASES.items()):
        kw = dict(kw)
        H = kw.pop("H")
        b = se.make_ekf_frames(1, 4100 + ci, **kw)
        fr = se.frame(b, 0)
        sel = se.make_selections(fr.ic, H, 4200 + ci)
        r = rn.ransac_hypotheses(fr, sel)
        for k in ("x", "P", "type", "pos", "has_z", "ic", "li0", "z", "h", "Hcam", "Hfeat", "R"):
            out[f"{name}_{k}"] = getattr(fr, k)
        out[f"{name}_std_z"] = fr.std_z
        out[f"{name}_sel"] = sel
        out[f"{name}_li"] = r["li"]
        out[f"{name}_supports"] = r["supports"]
        out[f"{name}_stats"] = np.array([r["n_hyp"], r["max_support"], r["best_hyp"], r["n_evaluated"], r["m"], r["num_ic"]])
        # hypothesised state + residuals of the first selection (tolerance checks)
        xi = rn.hypothesis_state(fr, list(sel[0][: r["m"]]))
        pattern, z_id, z_euc = rn.generate_state_vector_pattern(fr.type, fr.has_z, fr.z, fr.n)
        sup, li_id, li_euc, res = rn.compute_hypothesis_support_fast(xi, fr.cam, pattern, z_id, z_euc, fr.std_z, True)
        out[f"{name}_xi0"] = xi
        out[f"{name}_res0"] = res
        out[f"{name}_sup0"] = sup
        print(name, "n =", fr.n, "supports[:8] =", r["supports"][:8], "n_eval =", r["n_evaluated"], "max =", r["max_support"])
    np.savez_compressed(os.path.join(HERE, "ekf_frames.npz"), **out)


if __name__ == "__main__":
    main()

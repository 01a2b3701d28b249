"""Golden fixture from the reference's REAL descriptors: M/sift/data/box.sift and circle.sift (Lowe's text format:
"K 128" header, then per keypoint 4 frame numbers + 128 integers 0..255), matched by the REFERENCE's own siftmatch.c
(oracle/_ref, driven through its mexFunction by oracle/refmex.py).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden_box.py

Stored: the uint8 descriptors and, per case, the reference's 1-based matches and scores.  Cases: box vs itself,
box vs a perturbed + permuted copy, box vs circle; each as uint8 (sift_demo2.m:93-96) and as double(descr)/512
(class double holding float32 values, siftdescriptor.c:520-527)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refmex  # noqa: E402

DATA = "/root/reference/matlab_code/sift/data"


def read_sift(path):
    tok = open(path).read().split()
    K, nd = int(tok[0]), int(tok[1])
    v = np.array(tok[2:], dtype=np.float64).reshape(K, 4 + nd)
    return v[:, :4].copy(), v[:, 4:].astype(np.uint8)


def main():
    assert refmex.available()
    rng = np.random.Generator(np.random.PCG64(638))
    _, box = read_sift(os.path.join(DATA, "box.sift"))
    _, circle = read_sift(os.path.join(DATA, "circle.sift"))
    perm = rng.permutation(len(box))
    pert = np.clip(box[perm].astype(np.int32) + rng.integers(-6, 7, size=box.shape), 0, 255).astype(np.uint8)
    out = {"box": box, "circle": circle, "pert": pert, "perm": perm.astype(np.int32)}
    for name, (a, b) in {"self": (box, box), "pert": (box, pert), "circle": (box, circle)}.items():
        for cls in ("u8", "f64"):
            if cls == "u8":
                L1, L2 = a, b
            else:
                L1 = (a.astype(np.float32) / np.float32(512)).astype(np.float64)
                L2 = (b.astype(np.float32) / np.float32(512)).astype(np.float64)
            m, D = refmex.siftmatch(L1, L2, 1.5, nout=2)
            out[f"{name}_{cls}_matches"] = m
            out[f"{name}_{cls}_D"] = D
            print(name, cls, m.shape)
    np.savez_compressed(os.path.join(HERE, "box_sift.npz"), **out)


if __name__ == "__main__":
    main()

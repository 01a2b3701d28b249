"""Generates the committed golden fixtures under tests/golden/.

Run in the BUILD container (needs /root/reference): the siftmatch fixtures are produced by the
REFERENCE's own matlab_code/sift/siftmatch.c, compiled where it lies into
oracle/_ref/libsiftmatch_ref.so (oracle/Makefile) and driven through its real mexFunction
gateway (oracle/refmex.py).  Inputs are seeded; inputs and reference outputs are stored
together so the tests need nothing from /root/reference at run time.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refmex  # noqa: E402


def unit_desc(rng, K, nd=128):
    raw = np.abs(rng.normal(size=(K, nd)))
    d = raw / np.linalg.norm(raw, axis=1, keepdims=True)
    d = np.minimum(d, 0.2)
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    return d.astype(np.float32)


def case(name, L1, L2, thresh):
    m, D = refmex.siftmatch(L1, L2, thresh, nout=2)
    return {f"{name}_L1": L1, f"{name}_L2": L2, f"{name}_thresh": np.array(1.5 if thresh is None else thresh),
            f"{name}_matches": m, f"{name}_D": D}


def main():
    assert refmex.available(), "reference siftmatch.c not built (needs /root/reference)"
    rng = np.random.Generator(np.random.PCG64(20261018))
    out = {}
    # double holding float32 values (the pipeline's class, siftdescriptor.c:520-527), planted matches
    a = unit_desc(rng, 40)
    b = unit_desc(rng, 48)
    b[5:25] = a[10:30] + rng.normal(scale=0.002, size=(20, 128)).astype(np.float32)
    out.update(case("f64", a.astype(np.float64), b.astype(np.float64), None))
    out.update(case("f32", a, b, 1.5))
    # uint8 = round(512*d) as sift_demo2.m:93-94, with exact duplicates (ties -> first index)
    a8 = np.clip(np.floor(512.0 * a + 0.5), 0, 255).astype(np.uint8)
    b8 = np.clip(np.floor(512.0 * b + 0.5), 0, 255).astype(np.uint8)
    b8[30] = b8[6]
    b8[31] = b8[6]
    out.update(case("u8", a8, b8, 1.2))
    out.update(case("i8", (a8 // 2).astype(np.int8) - 20, (b8 // 2).astype(np.int8) - 20, 1.5))
    # odd ND, K2 = 1 (second_best stays at the start value -> always accepted), thresh = 1.0 with duplicates
    out.update(case("k2one", rng.normal(size=(7, 5)), rng.normal(size=(1, 5)), None))
    c = rng.normal(size=(9, 3))
    d = np.concatenate([c[[4, 4, 2]], rng.normal(size=(6, 3))])
    out.update(case("ties", c, d, 1.0))
    # the NN known answer of M/kNearestNeighbors.m:13-26 pushed through siftmatch: data 5x2, queries 3x2
    data = np.array([[1.0, 1.0], [2.0, 2.0], [3.0, 2.0], [4.0, 4.0], [5.0, 6.0]])
    query = np.array([[1.0, 1.0], [2.0, 1.0], [6.0, 2.0]])
    out.update(case("knn", query, data, 1.0))
    np.savez_compressed(os.path.join(HERE, "siftmatch_ref.npz"), **out)
    print("wrote siftmatch_ref.npz:", sorted(k for k in out if k.endswith("_matches")))
    for k in out:
        if k.endswith("_matches"):
            print(" ", k, out[k].shape)


if __name__ == "__main__":
    main()

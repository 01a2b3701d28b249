"""CPU suite, config 4: the 1-point-RANSAC EKF oracle (oracle/pre3_oracle_ekf.c).

"Parity unpinned": the reference ships no vectors for ransac_hypotheses /
compute_hypothesis_support_fast and cannot run here.  The C oracle is checked against the
independently written numpy/LAPACK restatement (oracle/ref_numpy_ekf.py): live, and through the
committed fixtures tests/golden/ekf_frames.npz (tests/golden/make_golden_ekf.py).
"""
import importlib
import math
import os

import numpy as np
import pytest

from oracle import ref_numpy_ekf as rn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["id_only", "mixed", "few_ic", "missing_z"]


@pytest.fixture(scope="module")
def se():
    return importlib.import_module("3pre_b200.synth_ekf")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "ekf_frames.npz"))


def golden_frame(se, gold, name):
    g = lambda k: gold[f"{name}_{k}"]
    x = g("x")
    return se.EkfFrame(n=len(x), F=len(g("type")), x=x, P=g("P"), type=g("type"), pos=g("pos"), has_z=g("has_z"),
                       ic=g("ic"), li0=g("li0"), z=g("z"), h=g("h"), Hcam=g("Hcam"), Hfeat=g("Hfeat"), R=g("R"),
                       cam=dict(se.CAM), std_z=float(g("std_z")), outlier=np.zeros(len(g("type")), bool))


def test_sincos_spec_within_one_ulp_of_libm(orc):
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-7, 7, 4000), rng.uniform(-0.8, 0.8, 1000), rng.uniform(-1e5, 1e5, 1000),
                         [0.0, math.pi / 4, -math.pi / 4, math.pi / 2, math.pi, 1e-300, 100.0]])
    for x in xs:
        s, c = orc.sincos(float(x))
        assert abs(s - math.sin(x)) <= np.spacing(abs(math.sin(x))) + 1e-300
        assert abs(c - math.cos(x)) <= np.spacing(abs(math.cos(x))) + 1e-300
    for bad in (float("nan"), float("inf"), 1e7):
        s, c = orc.sincos(bad)
        assert math.isnan(s) and math.isnan(c)


@pytest.mark.parametrize("name", CASES)
def test_oracle_vs_golden_restatement(orc, se, gold, name):
    fr = golden_frame(se, gold, name)
    sel = gold[f"{name}_sel"]
    n_hyp, max_support, best, n_eval, m, num_ic = gold[f"{name}_stats"]
    r = orc.ransac_hypotheses(fr, sel)
    assert (r["n_hyp"], r["max_support"], r["best_hyp"], r["n_evaluated"], r["m"], r["num_ic"]) == \
        (n_hyp, max_support, best, n_eval, m, num_ic)
    np.testing.assert_array_equal(r["supports"], gold[f"{name}_supports"])
    np.testing.assert_array_equal(r["li"], gold[f"{name}_li"])
    xi = orc.ekf_update(fr, sel[0][: int(m)])
    np.testing.assert_allclose(xi, gold[f"{name}_xi0"], rtol=0, atol=1e-9)
    pattern, z_id, z_euc = rn.generate_state_vector_pattern(fr.type, fr.has_z, fr.z, fr.n)
    sup, _, _, res = orc.ekf_support(xi, fr.cam, pattern, z_id.T, z_euc.T, fr.std_z)
    assert sup == int(gold[f"{name}_sup0"])
    np.testing.assert_allclose(res, gold[f"{name}_res0"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("kw", [dict(n_id=30, n_euc=0), dict(n_id=12, n_euc=12, interleave=True, asym=1e-6),
                                dict(n_id=20, n_euc=5, drop_z=0.2, drop_ic=0.2)])
def test_oracle_vs_numpy_live(orc, se, kw):
    b = se.make_ekf_frames(2, 4300 + len(kw), **kw)
    for f in range(2):
        fr = se.frame(b, f)
        sel = se.make_selections(fr.ic, 40, 77 + f)
        a, c = rn.ransac_hypotheses(fr, sel), orc.ransac_hypotheses(fr, sel)
        np.testing.assert_array_equal(a["supports"], c["supports"])
        np.testing.assert_array_equal(a["li"], c["li"])
        assert (a["n_hyp"], a["best_hyp"], a["n_evaluated"]) == (c["n_hyp"], c["best_hyp"], c["n_evaluated"])
        for i in range(3):
            np.testing.assert_allclose(orc.ekf_update(fr, sel[i][: a["m"]]),
                                       rn.hypothesis_state(fr, list(sel[i][: a["m"]])), rtol=0, atol=1e-9)


def test_planted_inliers_are_recovered(orc, se):
    """Hypotheses from three inlier matches support most inliers (the 1 px band sits on the smallest
    residual, :70, so noisy triples lose some) and never a gross outlier."""
    b = se.make_ekf_frames(1, 4400, n_id=60, outlier_ratio=0.25)
    fr = se.frame(b, 0)
    inl = np.flatnonzero(~fr.outlier)
    rng = np.random.default_rng(0)
    sel = np.stack([rng.permutation(inl)[:3] for _ in range(40)]).astype(np.int32)
    r = orc.ransac_hypotheses(fr, sel)
    assert r["max_support"] >= 0.8 * len(inl)
    assert not r["li"][fr.outlier].any()


def test_loop_semantics(orc, se):
    """n_hyp <= i at ransac_hypotheses.m:80 compares with the INNER loop variable (= m, :57): the loop
    runs on until n_hyp <= m, whatever the hypothesis counter; strict > keeps the first maximum."""
    b = se.make_ekf_frames(1, 4500, n_id=40, outlier_ratio=0.5)
    fr = se.frame(b, 0)
    out, inl = np.flatnonzero(fr.outlier), np.flatnonzero(~fr.outlier)
    bad = np.stack([out[:3], out[3:6], out[1:4]])
    good = np.stack([inl[:3], inl[3:6]])
    sel = np.concatenate([bad, good, good]).astype(np.int32)
    r = orc.ransac_hypotheses(fr, sel)
    # about half the matches are inliers: n_hyp = ceil(log(0.01)/log(eps)) stays > 3, no early stop
    assert r["n_evaluated"] == len(sel) and r["n_hyp"] > 3
    assert r["best_hyp"] == int(np.flatnonzero(r["supports"] == r["supports"].max())[0])
    # every match an inlier -> eps ~ 0 -> n_hyp <= 3 right after the first good hypothesis
    b2 = se.make_ekf_frames(1, 4501, n_id=40, outlier_ratio=0.0, meas_noise=0.05)
    fr2 = se.frame(b2, 0)
    r2 = orc.ransac_hypotheses(fr2, se.make_selections(fr2.ic, 30, 5))
    assert r2["n_evaluated"] < 30 and r2["n_hyp"] <= 3
    # the loop never runs past n_hyp_init
    r3 = orc.ransac_hypotheses(fr, sel, n_hyp_init=2)
    assert r3["n_evaluated"] == 2


def test_nhyp_rule(orc):
    assert orc.ekf_nhyp(50, 100) == math.ceil(math.log(1 - 0.99) / math.log(1 - (1 - (1 - 50 / 100))))
    assert orc.ekf_nhyp(100, 100) == 0.0                      # log(0) = -Inf
    assert orc.ekf_nhyp(80, 100) == 3.0 and orc.ekf_nhyp(78, 100) == 4.0
    for s, n in [(120, 100), (7, 5)]:                          # support > num_IC: complex log, real part
        assert orc.ekf_nhyp(s, n) == rn.n_hyp_rule(s, n)


def test_seeded_selection_is_a_permutation_prefix(orc):
    for num_ic in (1, 2, 3, 4, 7, 200):
        m = 3 if num_ic > 3 else 1
        seen = set()
        for hyp in range(300):
            r = orc.ekf_select(11, 3, hyp, num_ic, m)
            assert len(set(r.tolist())) == m and r.min() >= 0 and r.max() < num_ic
            seen.add(tuple(r.tolist()))
        if num_ic > 3:
            assert len(seen) > 20
            assert any(a > b for a, b, _ in seen)  # permutation order, not sorted


def test_no_ic_match_is_an_error(orc, se):
    b = se.make_ekf_frames(1, 4600, n_id=8, drop_ic=1.0)
    r = orc.ransac_hypotheses(se.frame(b, 0), None, H=10)
    assert r["status"] == 1 and r["n_evaluated"] == 0

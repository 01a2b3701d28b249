"""GPU parity, config 4: ransac_hypotheses / compute_hypothesis_support_fast through the C ABI
(libpre3.so) against the CPU oracle on the same seeded inputs -- supports, inlier masks and the
selected hypothesis bit-exact; states within 1e-9 of the independent numpy restatement."""
import importlib
import os

import numpy as np
import pytest

from oracle import ref_numpy_ekf as rn

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def se():
    return importlib.import_module("3pre_b200.synth_ekf")


def check_frames(ctx, orc, se, pre3, b, H, adaptive=True, supplied=True, seed=5, n_hyp_init=1000):
    Fr = b["x"].shape[0]
    frames = se.batch_to_numpy(b)
    sel = np.stack([se.make_selections(frames["ic"][f], H, 900 + f) for f in range(Fr)]) if supplied else None
    o = pre3.make_ekf_opts(n_hyp_init=n_hyp_init, H=H, adaptive=adaptive, seed=seed)
    res, li, sup = ctx.ransac_hypotheses_batch(frames, sel, o, frame_id0=40, want_supports=True)
    for f in range(Fr):
        fr = se.frame(b, f)
        r = orc.ransac_hypotheses(fr, sel[f] if supplied else None, H=H, n_hyp_init=n_hyp_init, seed=seed,
                                  frame_id=40 + f, adaptive=adaptive)
        assert res["status"][f] == r["status"]
        assert res["num_ic"][f] == r["num_ic"]
        if r["status"] != 0:
            continue
        assert (res["n_evaluated"][f], res["best_hyp"][f], res["max_support"][f], res["m"][f], res["n_hyp"][f]) == \
            (r["n_evaluated"], r["best_hyp"], r["max_support"], r["m"], r["n_hyp"])
        np.testing.assert_array_equal(sup[f, : r["n_evaluated"]], r["supports"])
        assert (sup[f, r["n_evaluated"]:] == -1).all()
        np.testing.assert_array_equal(li[f], r["li"])
    return res, li, sup


@pytest.mark.parametrize("kw", [dict(n_id=30, n_euc=0), dict(n_id=16, n_euc=12, interleave=True, asym=1e-6),
                                dict(n_id=24, n_euc=6, drop_z=0.2, drop_ic=0.2), dict(n_id=3, n_euc=0),
                                dict(n_id=70, n_euc=0, outlier_ratio=0.5)])
@pytest.mark.parametrize("supplied", [True, False])
def test_ransac_hypotheses_vs_oracle(ctx, orc, se, pre3, kw, supplied):
    b = se.make_ekf_frames(3, 4700 + kw["n_id"], **kw)
    check_frames(ctx, orc, se, pre3, b, H=80, supplied=supplied)


def test_ransac_hypotheses_waves_and_limits(ctx, orc, se, pre3):
    """More selections than one wave (32, 96, 224, ...): frames that stop early skip the later waves,
    frames with 50 % outliers run on; n_hyp_init caps the loop."""
    b = se.make_ekf_frames(4, 4800, n_id=40, outlier_ratio=0.5)
    res, _, _ = check_frames(ctx, orc, se, pre3, b, H=300)
    assert res["n_evaluated"].max() > 96
    b2 = se.make_ekf_frames(4, 4801, n_id=40, outlier_ratio=0.05)
    res2, _, _ = check_frames(ctx, orc, se, pre3, b2, H=300)
    assert res2["n_evaluated"].max() < 32
    check_frames(ctx, orc, se, pre3, b, H=300, n_hyp_init=50)
    check_frames(ctx, orc, se, pre3, b2, H=120, adaptive=False)
    check_frames(ctx, orc, se, pre3, b2, H=40, adaptive=False, supplied=False)


def test_ransac_hypotheses_golden(ctx, orc, se, pre3):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ekf_frames.npz"))
    for name in ["id_only", "mixed", "few_ic", "missing_z"]:
        g = lambda k: gold[f"{name}_{k}"]
        frames = {k: g(k)[None] for k in ("x", "type", "pos", "has_z", "ic", "li0", "z", "h", "Hcam", "Hfeat", "R")}
        frames["P"] = np.ascontiguousarray(g("P").T)[None]
        frames["std_z"], frames["cam"] = float(g("std_z")), dict(se.CAM)
        sel = g("sel")[None]
        o = pre3.make_ekf_opts(H=sel.shape[1])
        res, li, sup = ctx.ransac_hypotheses_batch(frames, sel, o, want_supports=True)
        n_hyp, max_support, best, n_eval, m, num_ic = g("stats")
        assert (res["n_hyp"][0], res["max_support"][0], res["best_hyp"][0], res["n_evaluated"][0], res["m"][0],
                res["num_ic"][0]) == (n_hyp, max_support, best, n_eval, m, num_ic)
        np.testing.assert_array_equal(sup[0, : int(n_eval)], g("supports"))
        np.testing.assert_array_equal(li[0], g("li"))


def test_ekf_edge_cases(ctx, orc, se, pre3):
    # no individually compatible match: select_random_match errors in the reference -> status 1, li untouched
    b = se.make_ekf_frames(2, 4900, n_id=8, drop_ic=1.0)
    frames = se.batch_to_numpy(b)
    frames["li0"][:] = 1
    res, li, _ = ctx.ransac_hypotheses_batch(frames, None, pre3.make_ekf_opts(H=10))
    assert (res["status"] == 1).all() and (li == 1).all()
    # H = 0: nothing evaluated, n_hyp stays at its initial value
    b = se.make_ekf_frames(1, 4901, n_id=8)
    res, li, _ = ctx.ransac_hypotheses_batch(se.batch_to_numpy(b), None, pre3.make_ekf_opts(H=0))
    assert res["status"][0] == 0 and res["n_evaluated"][0] == 0 and res["best_hyp"][0] == -1 and res["n_hyp"][0] == 1000
    # an IC feature without a measurement is rejected
    frames = se.batch_to_numpy(se.make_ekf_frames(1, 4902, n_id=8))
    frames["has_z"][0, 2] = 0
    res, _, _ = ctx.ransac_hypotheses_batch(frames, None, pre3.make_ekf_opts(H=4))
    assert res["status"][0] == 3
    # zero frames
    empty = {k: (v[:0] if isinstance(v, np.ndarray) else v) for k, v in frames.items()}
    res, li, _ = ctx.ransac_hypotheses_batch(empty, None, pre3.make_ekf_opts(H=4))
    assert len(res) == 0


def test_ekf_support_vs_oracle(ctx, orc, se):
    """compute_hypothesis_support_fast on given states: supports and masks bit-exact; also the
    states with a NaN / a point on the camera plane (residual NaN or Inf -> never an inlier)."""
    b = se.make_ekf_frames(1, 5000, n_id=40, n_euc=15, interleave=True, drop_z=0.15)
    fr = se.frame(b, 0)
    pattern, z_id, z_euc = rn.generate_state_vector_pattern(fr.type, fr.has_z, fr.z, fr.n)
    rng = np.random.default_rng(1)
    states = [fr.x] + [fr.x + 1e-3 * rng.normal(size=fr.n) for _ in range(6)]
    bad = fr.x.copy()
    bad[int(np.flatnonzero(pattern[:, 2])[0])] = np.nan
    states.append(bad)
    xi = np.stack(states)
    sup, li, le = ctx.ekf_support(xi, fr.cam, pattern, z_id.T, z_euc.T, fr.std_z)
    for k in range(len(states)):
        s, a, e, _ = orc.ekf_support(xi[k], fr.cam, pattern, z_id.T, z_euc.T, fr.std_z)
        assert sup[k] == s
        np.testing.assert_array_equal(li[k], a)
        np.testing.assert_array_equal(le[k], e)
        s2, a2, e2 = rn.compute_hypothesis_support_fast(xi[k], fr.cam, pattern, z_id, z_euc, fr.std_z)
        assert abs(int(s2) - int(s)) <= 1  # the numpy restatement differs only for residuals on the threshold
    # only cartesian / only inverse-depth measurements, and a pattern that does not fit
    s, a, e = ctx.ekf_support(fr.x, fr.cam, pattern * np.array([0, 0, 0, 1.0]), np.zeros((0, 2)), z_euc.T, fr.std_z)
    assert a.shape == (1, 0) and s[0] == orc.ekf_support(fr.x, fr.cam, pattern * np.array([0, 0, 0, 1.0]),
                                                         np.zeros((0, 2)), z_euc.T, fr.std_z)[0]
    with pytest.raises(Exception):
        ctx.ekf_support(fr.x, fr.cam, pattern, z_id.T[:-1], z_euc.T, fr.std_z)


def test_ekf_matlab_mirror(ctx, orc, se):
    ml = importlib.import_module("3pre_b200.matlab")
    ml._ctx = ctx
    b = se.make_ekf_frames(1, 5100, n_id=20, n_euc=5, drop_z=0.1)
    fr = se.frame(b, 0)
    filt, feats, cam = se.to_features_info(fr)
    sel = se.make_selections(fr.ic, 50, 3)
    out = ml.ransac_hypotheses(filt, feats, cam, selections=(sel.T + 1).astype(float))
    r = orc.ransac_hypotheses(fr, sel)
    got = np.array([f["low_innovation_inlier"] for f in out])
    np.testing.assert_array_equal(got, r["li"])
    assert ml.StatData["RANSAC_ITER"] == r["n_hyp"] and ml.StatData["RANSAC_HYP_SUPPORT"] == r["max_support"]
    # compute_hypothesis_support_fast with MATLAB shapes
    pattern, z_id, z_euc = rn.generate_state_vector_pattern(fr.type, fr.has_z, fr.z, fr.n)
    s, a, e = ml.compute_hypothesis_support_fast(fr.x.reshape(-1, 1), cam, pattern, z_id, z_euc, fr.std_z)
    so, ao, eo, _ = orc.ekf_support(fr.x, fr.cam, pattern, z_id.T, z_euc.T, fr.std_z)
    assert s == so and np.array_equal(a, ao) and np.array_equal(e, eo)
    s, a, e = ml.compute_hypothesis_support_fast(fr.x, cam, pattern * 0, [], [], fr.std_z)
    assert s == 0 and a.size == 0 and e.size == 0
    # H with a non-zero outside the camera / own-feature blocks is refused
    feats[0]["H"][0, 40] = 1.0
    with pytest.raises(ml.MexError):
        ml.ransac_hypotheses(filt, feats, cam)
    ml._ctx = None


def test_ekf_config4_shape_properties(ctx, se, pre3):
    """Config-4 shape (200 inverse-depth features, n = 1213, 1000 selections per frame), device path:
    equals the host path; the winner's inliers contain no gross outlier and most planted inliers."""
    import torch
    b = se.make_ekf_frames(6, 5200, device="cuda", n_id=200, outlier_ratio=0.2)
    assert b["n"] == 1213
    b["cam"] = dict(se.CAM)
    o = pre3.make_ekf_opts(H=1000, adaptive=False, seed=9)
    li = b["li0"].clone()
    res = torch.zeros(6, 32, dtype=torch.uint8, device="cuda")
    sup = torch.zeros(6, 1000, dtype=torch.int32, device="cuda")
    ctx.use_torch_stream()
    ctx.ransac_hypotheses_batch_dev(b, o, li, res, supports=sup)
    ctx.sync()
    r = np.frombuffer(res.cpu().numpy().tobytes(), dtype=pre3.EKF_RESULT_DTYPE)
    assert (r["status"] == 0).all() and (r["n_evaluated"] == 1000).all()
    hres, hli, hsup = ctx.ransac_hypotheses_batch(se.batch_to_numpy(b), None, o, want_supports=True)
    np.testing.assert_array_equal(hsup, sup.cpu().numpy())
    np.testing.assert_array_equal(hli, li.cpu().numpy())
    outl = b["outlier"].cpu().numpy()
    lin = li.cpu().numpy().astype(bool)
    assert not (lin & outl).any()
    assert (lin.sum(1) >= 0.8 * (~outl).sum(1)).all()
    assert (r["max_support"] == lin.sum(1)).all()
    # adaptive: a strict prefix of the same supports
    oa = pre3.make_ekf_opts(H=1000, adaptive=True, seed=9)
    ares, ali, asup = ctx.ransac_hypotheses_batch(se.batch_to_numpy(b), None, oa, want_supports=True)
    for f in range(6):
        n = ares["n_evaluated"][f]
        np.testing.assert_array_equal(asup[f, :n], hsup[f, :n])

"""GPU parity at the BASELINE.json shapes (-m gpu): every config's full-size shape against the oracle --
cfg2 2048 x 2048 x 128 (double and uint8) against the REFERENCE's own siftmatch.c (oracle/_ref) and the real
descriptors of M/sift/data/box.sift (committed golden, tests/golden/make_golden_box.py); cfg3 a 64-pair sequence
against orc.pair; cfg4 n = 1213 / 200 features / 1000 hypotheses against orc.ransac_hypotheses; cfg5 N = 20000
against orc.ransac on a 2000-hypothesis prefix (counts, states, full selection rule).  Plus the six MEX gateways
that were compile-only, driven through mexFunction (tests/mexdrv.py)."""
import importlib
import os

import numpy as np
import pytest

from oracle import ref_numpy as rn
from oracle import refmex

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-9


@pytest.fixture(scope="module")
def se():
    return importlib.import_module("3pre_b200.synth_ekf")


# ------------------------------------------------------------------------------------------
# cfg2: 2k x 2k x 128, class double and uint8, against the reference's compiled siftmatch.c
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cls", ["f64", "u8"])
@pytest.mark.parametrize("engine", [0, 1])
def test_cfg2_full_size_vs_reference(ctx, orc, synth, cls, engine):
    P, K = 2, 2048
    L1 = np.zeros((P, K, 128)); L2 = np.zeros((P, K, 128))
    for p in range(P):
        fp = synth.make_frame_pair(2000 + p, K1=K, K2=K, n_corr=K // 2)
        L1[p], L2[p] = fp.desc1, fp.desc2
    L2[0, 2047] = L2[0, 5]   # exact duplicate column: first index wins
    L1[1, 17] = L2[1, 40]    # an exact hit (distance 0)
    if cls == "u8":
        L1, L2 = synth.to_uint8(L1.reshape(-1, 128)).reshape(P, K, 128), synth.to_uint8(L2.reshape(-1, 128)).reshape(P, K, 128)
    ctx.set_match_engine(engine)
    try:
        out = ctx.siftmatch_batch(L1, L2, 1.5)
        one = ctx.siftmatch(L1[0], L2[0], 1.5, want_score=False)[0]
    finally:
        ctx.set_match_engine(0)
    for p in range(P):
        if refmex.available():   # the reference's own mexFunction
            m, D = refmex.siftmatch(L1[p], L2[p], 1.5, nout=2)
            ref_pairs, ref_D = (m.T - 1).astype(np.int32), D
        else:
            ref_pairs, ref_D = orc.siftmatch(L1[p], L2[p], 1.5)
        assert len(ref_pairs) >= 900
        np.testing.assert_array_equal(out[p][0], ref_pairs)
        np.testing.assert_array_equal(out[p][1], ref_D)
        if p == 0:
            np.testing.assert_array_equal(one, ref_pairs)


@pytest.mark.parametrize("engine", [0, 1])
def test_box_sift_real_descriptors(ctx, engine):
    """The reference's own fixture M/sift/data/box.sift (638 real uint8 descriptors): self match, perturbed +
    permuted copy, and box vs circle (no match survives the ratio test), as uint8 and as double / 512."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "box_sift.npz"))
    ctx.set_match_engine(engine)
    try:
        for name, (a, b) in {"self": ("box", "box"), "pert": ("box", "pert"), "circle": ("box", "circle")}.items():
            for cls in ("u8", "f64"):
                L1, L2 = g[a], g[b]
                if cls == "f64":
                    L1 = (L1.astype(np.float32) / np.float32(512)).astype(np.float64)
                    L2 = (L2.astype(np.float32) / np.float32(512)).astype(np.float64)
                pairs, score = ctx.siftmatch(L1, L2, 1.5)
                np.testing.assert_array_equal(pairs.T + 1, g[f"{name}_{cls}_matches"].astype(np.int64))
                np.testing.assert_array_equal(score, g[f"{name}_{cls}_D"])
        # the perturbed copy is found where the permutation put it (for the rows the reference accepts)
        m = g["pert_u8_matches"].astype(np.int64) - 1
        assert m.shape[1] > 500 and (g["perm"][m[1]] == m[0]).mean() > 0.99
    finally:
        ctx.set_match_engine(0)


# ------------------------------------------------------------------------------------------
# cfg3: a 64-pair sequence (K = 512, 300 re-observed features, H = 2000, adaptive) against orc.pair
# ------------------------------------------------------------------------------------------
def test_cfg3_sequence_64_pairs_vs_oracle(ctx, orc, synth, pre3):
    F = 65
    sq = synth.make_sequence_torch(F, 3003, "cuda", K=512, n_corr=300)
    desc, xyz = sq["desc"].cpu().numpy(), sq["xyz"].cpu().numpy()
    opts = pre3.make_opts(method=0, k=5, max_iteration=2000, adaptive=True, H=2000, seed=11)
    res, matches, masks = ctx.sequence(desc, xyz, opts, pair_id0=100)
    assert len(res) == F - 1
    for p in range(F - 1):
        om, o = orc.pair(desc[p], desc[p + 1], xyz[p], xyz[p + 1], 11, 100 + p, H=2000)
        n = int(res["n_matches"][p])
        assert n == len(om) and n >= 250
        np.testing.assert_array_equal(matches[p, :n], om)
        g = pre3.unpack_result(res[p], masks[p, :n].astype(bool))
        assert (g.status, g.state, g.best_fit, g.best_sample, g.best_iter, g.n_iter, g.n_consumed) == \
            (o.status, o.state, o.best_fit, o.best_sample, o.best_iter, o.n_iter, o.n_consumed), p
        assert g.thr == o.thr and g.error_sum == o.error_sum
        np.testing.assert_array_equal(g.mask, o.mask)
        np.testing.assert_array_equal(g.R_hyp, o.R_hyp)
        np.testing.assert_array_equal(g.T_hyp, o.T_hyp)
        assert rn.rot_angle(g.R, o.R) < TOL and np.abs(g.T - o.T).max() < TOL
        assert rn.rot_angle(g.R, sq["R"][p].cpu().numpy()) < 3e-3


# ------------------------------------------------------------------------------------------
# cfg4: n = 1213 (200 inverse-depth features), 1000 hypotheses per frame, against orc.ransac_hypotheses
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("adaptive", [False, True])
def test_cfg4_full_size_vs_oracle(ctx, orc, se, pre3, adaptive):
    from test_gpu_ekf import check_frames
    b = se.make_ekf_frames(2, 5300, n_id=200, outlier_ratio=0.2)
    assert b["n"] == 1213
    res, li, sup = check_frames(ctx, orc, se, pre3, b, H=1000, adaptive=adaptive, supplied=True)
    if not adaptive:
        assert (res["n_evaluated"] == 1000).all()
    res2, _, _ = check_frames(ctx, orc, se, pre3, b, H=1000, adaptive=adaptive, supplied=False, seed=21)
    assert (res2["max_support"] >= 120).all()


# ------------------------------------------------------------------------------------------
# cfg5: N = 20000, 60 % outliers: the first 2000 hypotheses of the 1M run against orc.ransac
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [5, 3])
def test_cfg5_prefix_vs_oracle(ctx, orc, synth, pre3, k):
    N, Hp = 20000, 2000
    c = synth.make_correspondences(5005, N=N, outlier_ratio=0.60)
    samples = orc.sample_sets(1, 0, Hp, N, k)       # the seeded generator's first Hp sets of pair 0
    o = orc.ransac(c.Ya, c.Yb, samples, method=0, max_iteration=Hp + 1, adaptive=False)
    # supplied sets
    g = ctx.ransac(c.Ya, c.Yb, samples, pre3.make_opts(method=0, k=k, max_iteration=Hp + 1, adaptive=False, H=Hp))
    # seeded: the same sets come out of the device generator
    g2 = ctx.ransac(c.Ya, c.Yb, None, pre3.make_opts(method=0, k=k, max_iteration=Hp + 1, adaptive=False, H=Hp, seed=1))
    for r in (g, g2):
        np.testing.assert_array_equal(r.states, o.states)
        np.testing.assert_array_equal(r.counts, o.counts)
        assert (r.status, r.state, r.best_fit, r.best_sample, r.best_iter, r.n_iter, r.n_consumed) == \
            (o.status, o.state, o.best_fit, o.best_sample, o.best_iter, o.n_iter, o.n_consumed)
        assert r.thr == o.thr and r.error_sum == o.error_sum
        np.testing.assert_array_equal(r.mask, o.mask)
        np.testing.assert_array_equal(r.R_hyp, o.R_hyp)
        assert rn.rot_angle(r.R, o.R) < TOL and np.abs(r.T - o.T).max() < TOL
    # the 1M-hypothesis run's first block is this prefix: a 100k run reports the same first 2000 cardinalities
    g3 = ctx.ransac(c.Ya, c.Yb, None, pre3.make_opts(method=0, k=k, max_iteration=100001, adaptive=False, H=100000, seed=1))
    np.testing.assert_array_equal(g3.counts[:Hp], o.counts)
    np.testing.assert_array_equal(g3.states[:Hp], o.states)


# ------------------------------------------------------------------------------------------
# the six MEX gateways that were compile-only: struct `options`, `cam`, logical outputs, struct outputs
# ------------------------------------------------------------------------------------------
def test_gateway_RANSAC_CALC_VER2_mex(orc, synth):
    from mexdrv import Gateway
    gw = Gateway("RANSAC_CALC_VER2_mex")
    c = synth.make_correspondences(8, N=200, outlier_ratio=0.3)
    samples = synth.make_samples(9, 800, 200, 5)
    options = {"DistanceThreshold": 0.05, "MaxIteration": 2000}
    R, T, err, best_fit, state = gw(c.Ya.T, c.Yb.T, options, (samples.T + 1).astype(np.float64), nout=5)
    o = orc.ransac(c.Ya, c.Yb, samples, method=0, max_iteration=2000)
    assert best_fit[0, 0] == o.best_fit and state[0, 0] == o.state and T.shape == (3, 1) and R.shape == (3, 3)
    assert set(err) == {"ErrorSum", "mYa", "mYb"} and err["ErrorSum"][0, 0] == o.error_sum
    np.testing.assert_array_equal(err["mYa"], c.Ya[o.mask].T)
    np.testing.assert_array_equal(err["mYb"], c.Yb[o.mask].T)
    assert rn.rot_angle(R, o.R) < TOL and np.abs(T.ravel() - o.T).max() < TOL
    # seeded form + k + adaptive off + Horn
    R2, T2, _, bf2, _ = gw(c.Ya.T, c.Yb.T, {"DistanceThreshold": 0.02, "MaxIteration": 300}, 7.0, 4.0, 0.0, 1.0, nout=5)
    o2 = orc.ransac(c.Ya, c.Yb, orc.sample_sets(7, 0, 300, 200, 4), method=1, max_iteration=300, distance_threshold=0.02,
                    adaptive=False)
    assert bf2[0, 0] == o2.best_fit and rn.rot_angle(R2, o2.R) < TOL
    with pytest.raises(refmex.MexError, match="struct"):
        gw(c.Ya.T, c.Yb.T, 5.0, nout=1)
    with pytest.raises(refmex.MexError, match="DistanceThreshold"):
        gw(c.Ya.T, c.Yb.T, {"MaxIteration": 10}, nout=1)
    with pytest.raises(refmex.MexError, match="same size"):
        gw(c.Ya.T, c.Yb[:50].T, options, nout=1)
    with pytest.raises(refmex.MexError, match="get_rand"):   # fewer correspondences than k
        gw(c.Ya[:3].T, c.Yb[:3].T, options, nout=1)


def test_gateway_find_transform_matrix_and_horn(orc):
    from mexdrv import Gateway
    rng = np.random.default_rng(2)
    A = rng.normal(size=(40, 3))
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    B = A @ R.T * 1.7 + np.array([0.3, -0.2, 0.9]) + 1e-3 * rng.normal(size=A.shape)
    ftm = Gateway("find_transform_matrix_mex")
    rot, trans, state = ftm(B.T, A.T, nout=3)
    r0, t0, s0 = orc.find_transform_matrix(B, A)
    assert state[0, 0] == s0
    np.testing.assert_array_equal(rot, r0)
    np.testing.assert_array_equal(trans.ravel(), t0)
    with pytest.raises(refmex.MexError, match="same size"):
        ftm(B.T, A[:5].T, nout=1)
    horn = Gateway("absoluteOrientationQuaternion_mex")
    s, Rh, Th, e = horn(A.T, B.T, nout=4)                      # doScale defaults to 1 (:32-34)
    s0, R0, T0, e0 = orc.horn(A, B, 1)
    assert s[0, 0] == s0 and e[0, 0] == e0 and abs(s0 - 1.7) < 1e-2
    np.testing.assert_array_equal(Rh, R0)
    np.testing.assert_array_equal(Th.ravel(), T0)
    s1, R1, T1, e1 = horn(A.T, B.T, 0.0, nout=4)
    s2, R2, T2, e2 = orc.horn(A, B, 0)
    assert s1[0, 0] == 1.0 and e1[0, 0] == e2
    np.testing.assert_array_equal(T1.ravel(), T2)
    with pytest.raises(refmex.MexError, match="at least 4"):
        horn(A[:3].T, B[:3].T, nout=1)
    with pytest.raises(refmex.MexError, match="dimension 3"):
        horn(A[:, :2].T, B[:, :2].T, nout=1)


def test_gateway_compute_hypothesis_support_fast_mex(orc, se):
    from mexdrv import Gateway
    from oracle import ref_numpy_ekf as rne
    gw = Gateway("compute_hypothesis_support_fast_mex")
    b = se.make_ekf_frames(1, 5000, n_id=40, n_euc=15, interleave=True, drop_z=0.15)
    fr = se.frame(b, 0)
    pattern, z_id, z_euc = rne.generate_state_vector_pattern(fr.type, fr.has_z, fr.z, fr.n)
    cam = dict(se.CAM)
    s, li, le = gw(fr.x.reshape(-1, 1), cam, pattern, z_id, z_euc, fr.std_z, nout=3)
    s0, li0, le0, _ = orc.ekf_support(fr.x, cam, pattern, z_id.T, z_euc.T, fr.std_z)
    assert s[0, 0] == s0 and li.dtype == bool and li.shape == (1, z_id.shape[1]) and le.shape == (1, z_euc.shape[1])
    np.testing.assert_array_equal(li.ravel(), li0.astype(bool))
    np.testing.assert_array_equal(le.ravel(), le0.astype(bool))
    # several states as columns
    X = np.stack([fr.x, fr.x + 1e-3, fr.x * (1 + 1e-4)], 1)
    s3, li3, _ = gw(X, cam, pattern, z_id, z_euc, fr.std_z, nout=3)
    for j in range(3):
        sj, lij, _, _ = orc.ekf_support(X[:, j], cam, pattern, z_id.T, z_euc.T, fr.std_z)
        assert s3[0, j] == sj
        np.testing.assert_array_equal(li3[j], lij.astype(bool))
    # no euclidean features: empty z_euc -> [] mask (:112-116)
    b2 = se.make_ekf_frames(1, 5001, n_id=20, n_euc=0)
    f2 = se.frame(b2, 0)
    p2, zi2, ze2 = rne.generate_state_vector_pattern(f2.type, f2.has_z, f2.z, f2.n)
    s4, li4, le4 = gw(f2.x.reshape(-1, 1), cam, p2, zi2, np.zeros((0, 0)), f2.std_z, nout=3)
    s5, li5, _, _ = orc.ekf_support(f2.x, cam, p2, zi2.T, ze2.T, f2.std_z)
    assert le4.size == 0 and s4[0, 0] == s5
    np.testing.assert_array_equal(li4.ravel(), li5.astype(bool))
    # a pattern with euclidean features but no euclidean measurements is the reference's reshape error
    with pytest.raises(refmex.MexError, match="reshape"):
        gw(fr.x.reshape(-1, 1), cam, pattern, z_id, np.zeros((0, 0)), fr.std_z, nout=1)
    with pytest.raises(refmex.MexError, match="cam"):
        gw(fr.x.reshape(-1, 1), 1.0, pattern, z_id, z_euc, fr.std_z, nout=1)
    with pytest.raises(refmex.MexError, match="missing"):
        gw(fr.x.reshape(-1, 1), {"f": 1.0}, pattern, z_id, z_euc, fr.std_z, nout=1)


def test_gateway_ransac_hypotheses_mex(orc, se):
    from mexdrv import Gateway
    gw = Gateway("ransac_hypotheses_mex")
    b = se.make_ekf_frames(1, 5100, n_id=24, n_euc=6, interleave=True)
    fr = se.frame(b, 0)
    F = fr.F
    sel = se.make_selections(fr.ic, 64, 77)
    args = [fr.x.reshape(-1, 1), fr.P, fr.std_z, dict(se.CAM), fr.type.astype(float), (fr.pos + 1).astype(float),
            fr.has_z.astype(float), fr.ic.astype(float), fr.z.T, fr.h.T,
            np.transpose(fr.Hcam, (2, 1, 0)), np.transpose(fr.Hfeat, (2, 1, 0)), np.transpose(fr.R, (2, 1, 0)),
            fr.li0.astype(float)]
    li, stats = gw(*args, (sel.T + 1).astype(np.float64), nout=2)
    o = orc.ransac_hypotheses(fr, sel, H=64)
    np.testing.assert_array_equal(li.ravel().astype(np.uint8), o["li"])
    assert (stats[0, 1], stats[0, 2] - 1, stats[0, 3], stats[0, 4]) == (o["max_support"], o["best_hyp"], o["n_evaluated"],
                                                                      o["num_ic"])
    assert stats[0, 0] == o["n_hyp"]
    # seeded + n_hyp + adaptive off
    li2, stats2 = gw(*args, 3.0, 40.0, 0.0, nout=2)
    o2 = orc.ransac_hypotheses(fr, None, H=40, n_hyp_init=40, seed=3, frame_id=0, adaptive=False)
    np.testing.assert_array_equal(li2.ravel().astype(np.uint8), o2["li"])
    assert stats2[0, 3] == o2["n_evaluated"] == 40
    with pytest.raises(refmex.MexError, match="14 input"):
        gw(*args[:5], nout=1)
    bad = list(args); bad[8] = fr.z.T[:, :3]
    with pytest.raises(refmex.MexError, match="wrong class or size"):
        gw(*bad, nout=1)


def test_gateway_read_xyz_sr4000_mex(orc, synth):
    from mexdrv import Gateway
    gw = Gateway("read_xyz_sr4000_mex")
    sr, frames = synth.make_sr_frames(6100, 1, 60)
    srm = sr[0].T                                  # MATLAB rows x 176
    x, y, z, conf, mc = gw(srm, nout=5)
    ox, oy, oz = orc.read_xyz(sr[0])

    def eq_nan(a, b):
        np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
        np.testing.assert_array_equal(np.nan_to_num(a, nan=-7.0), np.nan_to_num(b, nan=-7.0))

    eq_nan(x, ox.T); eq_nan(y, oy.T); eq_nan(z, oz.T)
    np.testing.assert_array_equal(conf, srm[576:720])
    assert mc[0, 0] == orc.max_confidence(sr[0])
    xyz, idx = gw(srm, frames[0].T, 1.0, nout=2)  # any third argument selects the per-feature form
    oxyz, keep, oidx, _ = orc.features_xyz(sr[0], frames[0])
    np.testing.assert_array_equal(idx.ravel().astype(int) - 1, oidx)
    eq_nan(xyz, oxyz.T)
    assert keep.sum() == len(oidx) > 10
    with pytest.raises(refmex.MexError, match="176 columns"):
        gw(srm[:, :100], nout=1)

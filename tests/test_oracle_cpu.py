"""CPU suite, part 1: the oracle itself.

* siftmatch restatement vs the REFERENCE's siftmatch.c outputs (committed golden fixtures,
  tests/golden/make_golden.py) and, when oracle/_ref is present, vs the reference library
  live on fresh random inputs -> this stage of the oracle is PINNED;
* Kabsch / Horn / RANSAC restatement vs the independent LAPACK restatement
  (oracle/ref_numpy.py) to 1e-9 and vs planted ground truth -> "parity unpinned" stages
  (the reference ships no vectors for them).
"""
import importlib

import numpy as np
import pytest

from oracle import ref_numpy as rn
from oracle import refmex

CASES = ["f64", "f32", "u8", "i8", "k2one", "ties", "knn"]


@pytest.mark.parametrize("name", CASES)
def test_siftmatch_oracle_vs_reference_golden(orc, golden, name):
    L1, L2 = golden[f"{name}_L1"], golden[f"{name}_L2"]
    pairs, score = orc.siftmatch(L1, L2, float(golden[f"{name}_thresh"]))
    np.testing.assert_array_equal(pairs.T + 1, golden[f"{name}_matches"].astype(np.int64))
    np.testing.assert_array_equal(score, golden[f"{name}_D"])  # bit-exact scores


def test_knn_known_answer(orc, golden):
    # M/kNearestNeighbors.m:13-26: a = [1 1;2 2;3 2;4 4;5 6], b = [1 1;2 1;6 2] -> nearest ids
    # [1;1;4] (query 2 ties data 1 and 2 at distance 1.0 -> lower index first), distances
    # [0;1;2.8284]; siftmatch reports the squared distance
    np.testing.assert_array_equal(golden["knn_matches"][0], [1, 2, 3])
    np.testing.assert_array_equal(golden["knn_matches"][1], [1, 1, 4])
    np.testing.assert_allclose(np.sqrt(golden["knn_D"]), [0.0, 1.0, 2.8284], atol=5e-5)


@pytest.mark.skipif(not refmex.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.uint8, np.int8])
def test_siftmatch_oracle_vs_reference_live(orc, dtype):
    rng = np.random.default_rng(7)
    for trial in range(6):
        K1, K2, ND = rng.integers(1, 60), rng.integers(1, 60), int(rng.choice([1, 3, 64, 128]))
        if np.issubdtype(dtype, np.floating):
            L1 = rng.normal(size=(K1, ND)).astype(dtype)
            L2 = rng.normal(size=(K2, ND)).astype(dtype)
        else:
            info = np.iinfo(dtype)
            L1 = rng.integers(info.min, info.max + 1, size=(K1, ND)).astype(dtype)
            L2 = rng.integers(info.min, info.max + 1, size=(K2, ND)).astype(dtype)
        if K2 > 3:
            L2[K2 - 1] = L2[0]  # duplicate columns: first index must win
        thresh = float(rng.choice([1.0, 1.5, 1.01]))
        m, D = refmex.siftmatch(L1, L2, thresh, nout=2)
        pairs, score = orc.siftmatch(L1, L2, thresh)
        np.testing.assert_array_equal(pairs.T + 1, m.astype(np.int64))
        np.testing.assert_array_equal(score, D)


@pytest.mark.skipif(not refmex.available(), reason="oracle/_ref not built")
def test_reference_gateway_errors():
    a = np.zeros((3, 4))
    with pytest.raises(refmex.MexError, match="same number of rows"):
        refmex.siftmatch(a, np.zeros((3, 5)))
    with pytest.raises(refmex.MexError, match="same class"):
        refmex.siftmatch(a, a.astype(np.float32))
    with pytest.raises(refmex.MexError, match="Unsupported numeric class"):
        refmex.siftmatch(a.astype(np.int32), a.astype(np.int32))
    with pytest.raises(refmex.MexError, match="At most three"):
        refmex.siftmatch(a, a, 1.5, extra_args=1)


def _rand_rigid(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    return R, rng.normal(size=3)


@pytest.mark.parametrize("n", [3, 4, 5, 50, 1000])
def test_fits_vs_lapack(orc, n):
    rng = np.random.default_rng(100 + n)
    for _ in range(20):
        R, t = _rand_rigid(rng)
        Yb = rng.normal(size=(n, 3)) * 2
        Ya = Yb @ R.T + t + rng.normal(scale=1e-3, size=(n, 3))
        rot, tr, st = orc.find_transform_matrix(Ya, Yb)
        r2, t2, st2 = rn.find_transform_matrix(Ya, Yb)
        assert st == 1 and st2 in (1, 2)
        assert rn.rot_angle(rot, r2) < 1e-9 and np.abs(tr - t2).max() < 1e-9
        assert rn.rot_angle(rot, R) < 5e-2
        s, Rh, Th, err = orc.horn(Yb, Ya, doScale=0, allow_small=True)
        s2, R2, T2, err2 = rn.horn(Yb, Ya, False)
        assert rn.rot_angle(Rh, R2) < 1e-9 and np.abs(Th - T2).max() < 1e-9 and abs(err - err2) < 1e-9 * max(1, err2)
        # Kabsch and Horn minimise the same cost
        assert rn.rot_angle(Rh, rot) < 1e-7
        if n >= 4:
            s, Rh, Th, err = orc.horn(Yb, Ya, doScale=1)
            s2, R2, T2, err2 = rn.horn(Yb, Ya, True)
            assert abs(s - s2) < 1e-9 and np.abs(Th - T2).max() < 1e-9


def test_horn_needs_four_points(orc):
    with pytest.raises(ValueError, match="at least 4"):
        orc.horn(np.zeros((3, 3)), np.zeros((3, 3)))


def test_degenerate_fits(orc):
    # collinear sample: two vanishing singular values -> state -1, rot = H, trans = 0
    p = np.outer(np.arange(5.0), [1.0, 2.0, 3.0])
    rot, tr, st = orc.find_transform_matrix(p + 1.0, p)
    assert st == -1 and np.all(tr == 0)
    # coplanar (rank 2): a proper rotation is still recovered (state 1 by the oracle's convention)
    rng = np.random.default_rng(3)
    R, t = _rand_rigid(rng)
    Yb = np.c_[rng.normal(size=(6, 2)), np.zeros(6)]
    Ya = Yb @ R.T + t
    rot, tr, st = orc.find_transform_matrix(Ya, Yb)
    assert st == 1 and rn.rot_angle(rot, R) < 1e-9 and abs(np.linalg.det(rot) - 1) < 1e-12
    # non-finite input -> state 0
    bad = Ya.copy()
    bad[0, 0] = np.nan
    assert orc.find_transform_matrix(bad, Yb)[2] == 0


@pytest.mark.parametrize("method,k", [(0, 5), (0, 3), (1, 5), (1, 4)])
@pytest.mark.parametrize("adaptive", [True, False])
def test_ransac_oracle_vs_lapack_restatement(orc, method, k, adaptive):
    synth = importlib.import_module("3pre_b200.synth")
    for seed in range(4):
        c = synth.make_correspondences(1000 + seed, N=300, outlier_ratio=0.30)
        samples = synth.make_samples(2000 + seed, 400 if not adaptive else 2000, 300, k)
        r = orc.ransac(c.Ya, c.Yb, samples, method=method, max_iteration=2000, distance_threshold=0.012,
                       adaptive=adaptive)
        g = rn.ransac_ver2(c.Ya, c.Yb, samples, 2000, adaptive, method, 0.012)
        assert r.status == 0
        assert r.best_fit == g["best_fit"] and r.best_sample == g["best_sample"] and r.n_iter == g["n_iter"]
        np.testing.assert_array_equal(r.mask, g["mask"])
        assert abs(r.thr - g["thr"]) < 1e-15 and abs(r.error_sum - g["error_sum"]) < 1e-9
        assert rn.rot_angle(r.R, g["R"]) < 1e-9 and np.abs(r.T - g["T"]).max() < 1e-9
        # planted motion recovered, planted inliers found (noise 2 mm << thr ~ 8-15 mm << outliers)
        assert rn.rot_angle(r.R, c.R) < 2e-3 and np.abs(r.T - c.t).max() < 5e-3
        assert (r.mask & ~c.inlier).sum() <= 2 and (r.mask & c.inlier).sum() >= 0.9 * c.inlier.sum()


def test_ransac_loop_semantics(orc):
    synth = importlib.import_module("3pre_b200.synth")
    c = synth.make_correspondences(5, N=120, outlier_ratio=0.2)
    samples = synth.make_samples(6, 50, 120, 5)
    # while iter < MaxIteration runs at most MaxIteration-1 times (RANSAC_CALC_VER2.m:86)
    r = orc.ransac(c.Ya, c.Yb, samples, max_iteration=10, adaptive=False)
    # (samples that contain outliers can give a reflection, det = -1 -> state -1 -> skipped uncounted)
    assert r.n_iter == 9 and r.n_consumed == 9 + int((r.states[: r.n_consumed] == -1).sum())
    # fewer correspondences than k: get_rand errors -> status 1
    assert orc.ransac(c.Ya[:4], c.Yb[:4], samples % 4, adaptive=False).status == 1
    # perfect data: card == N -> nIterations = 0 -> stops right after the first hypothesis
    Ya = c.Yb @ c.R.T + c.t
    r = orc.ransac(Ya, c.Yb, samples, adaptive=True)
    assert r.n_iter == 1 and r.best_fit == 120 and r.best_sample == 0
    # a degenerate (collinear) first sample is skipped without counting (:97-99)
    Yb = c.Yb.copy()
    Yb[:5] = np.outer(np.arange(5.0), [1.0, 1.0, 1.0])
    Ya = Yb @ c.R.T + c.t
    s2 = samples.copy()
    s2[0] = np.arange(5)
    r = orc.ransac(Ya, Yb, s2, adaptive=False, max_iteration=20)
    nskip = int((r.states[: r.n_consumed] == -1).sum())
    assert r.states[0] == -1 and r.counts[0] == -1 and r.n_iter == 19 and r.n_consumed == 19 + nskip


def test_adaptive_rule_values(orc):
    # 5*ceil(log(0.01)/log(1-(c/N)^5)) (RANSAC_CALC_VER2.m:139)
    assert orc.adaptive_niter(300, 300) == 0
    assert orc.adaptive_niter(210, 300) == 5 * np.ceil(np.log(0.01) / np.log(1 - 0.7 ** 5))
    assert orc.adaptive_niter(0, 300) == -np.inf
    assert orc.adaptive_niter(150, 300, k=3, mult=1) == np.ceil(np.log(0.01) / np.log(1 - 0.5 ** 3))


def test_sample_sets_are_ascending_distinct(orc):
    s = orc.sample_sets(123, 7, 500, 37, 5)
    assert s.min() >= 0 and s.max() < 37
    assert np.all(np.diff(s, axis=1) > 0)
    # every index is reachable and the sets differ between pairs
    assert len(np.unique(s)) == 37
    assert not np.array_equal(s, orc.sample_sets(123, 8, 500, 37, 5))


def test_R2q(orc):
    rng = np.random.default_rng(0)
    for _ in range(10):
        R, _ = _rand_rigid(rng)
        q = orc.R2q(R)
        assert abs(np.linalg.norm(q) - 1) < 1e-12
        # q = [a -b -c -d]: conjugate convention of the slamToolbox (R2q.m:55); q2R(q) == R
        a, b, c, d = q
        Rq = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                       [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                       [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])
        assert np.abs(Rq - R).max() < 1e-12
    assert np.allclose(orc.R2q(np.diag([1.0, -1.0, -1.0])), [0, -1, 0, 0])


def test_matching_gate_known_answer(orc):
    """matching_sift_based.m:117-150 on a hand-made case: three predicted features with unambiguous matches; the
    radius comes from S of the i-th PREDICTED feature (loop counter), 40 when S is empty."""
    e = np.eye(8)
    des1 = np.stack([10 * e[0], 10 * e[1], 10 * e[2]])
    des2 = np.stack([10 * e[2], 10 * e[0], 10 * e[1], 10 * e[5]])          # matches: (0,1) (1,2) (2,0)
    pos2 = np.array([[100.0, 100.0], [10.0, 10.0], [50.0, 50.0], [0.0, 0.0]])
    h = np.array([[10.0, 49.0],      # dist to pos2[1] = 39 <= 40 (S empty)            -> compatible
                  [50.0, 56.5],      # dist to pos2[2] = 6.5 > ceil(3*sqrt(4)) = 6       -> discarded
                  [100.0, 91.0]])    # dist to pos2[0] = 9 <= ceil(3*sqrt(8.5)) = 9      -> compatible (<=)
    S11 = np.array([np.nan, 4.0, 8.5])
    ic, z, mt, nm, nd = orc.matching_sift_based(des1, des2, h, S11, pos2)
    assert nm == 3 and nd == 1
    np.testing.assert_array_equal(ic, [True, False, True])
    np.testing.assert_array_equal(mt, [1, -1, 0])
    np.testing.assert_array_equal(z[0], pos2[1])
    np.testing.assert_array_equal(z[2], pos2[0])
    assert np.isnan(z[1]).all()

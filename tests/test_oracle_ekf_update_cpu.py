"""CPU suite: the checker of the EKF partial updates (oracle/ref_numpy_ekf.py: update,
ekf_update_inliers, rescue_hi_inliers, normJac) -- the dense numpy / LAPACK restatement of M/update.m,
M/@ekf_filter/ekf_update_li_inliers.m, ekf_update_hi_inliers.m, rescue_hi_inliers.m.  PARITY UNPINNED (no reference
vectors): the restatement is checked against the information form of the same update and against finite differences."""
import importlib

import numpy as np

from oracle import ref_numpy_ekf as rne

se = importlib.import_module("3pre_b200.synth_ekf")


def test_normJac_is_the_jacobian_of_q_over_norm():
    rng = np.random.default_rng(0)
    q = rng.normal(size=4)
    J = rne.normJac(q)
    num = np.zeros((4, 4))
    for k in range(4):
        d = np.zeros(4); d[k] = 1e-6
        num[:, k] = ((q + d) / np.linalg.norm(q + d) - (q - d) / np.linalg.norm(q - d)) / 2e-6
    assert np.abs(J - num).max() < 1e-8          # M/normJac.m:1-16


def test_update_matches_the_information_form():
    b = se.make_ekf_frames(1, 31, n_id=8, n_euc=3, interleave=True)
    fr = se.frame(b, 0)
    flags = np.ones(fr.F, np.uint8); flags[2] = 0
    sel = [i for i in range(fr.F) if flags[i]]
    H = np.vstack([rne.dense_H(fr, i) for i in sel])
    z = np.concatenate([fr.z[i] for i in sel]); h = np.concatenate([fr.h[i] for i in sel])
    x, p, K = rne.update(fr.x, fr.P, H, np.eye(len(z)), z, h)
    # before the quaternion step: (P^-1 + H' H)^-1 and x + P' H' (z - h)
    Pi = np.linalg.inv(np.linalg.inv(fr.P) + H.T @ H)
    xi = fr.x + Pi @ H.T @ (z - h)
    J = rne.normJac(xi[3:7])
    D = np.eye(fr.n); D[3:7, 3:7] = J
    assert np.abs(p - D @ Pi @ D.T).max() < 1e-9 * np.abs(p).max()
    xi[3:7] /= np.linalg.norm(xi[3:7])
    assert np.abs(x - xi).max() < 1e-9
    assert np.abs(p - p.T).max() < 1e-18 + 1e-12 * np.abs(p).max() and abs(np.linalg.norm(x[3:7]) - 1) < 1e-15
    x2, p2 = rne.ekf_update_inliers(fr, flags)
    assert np.array_equal(x2, x) and np.array_equal(p2, p)
    # nothing flagged: copied through (M/update.m:50-54)
    x0, p0 = rne.ekf_update_inliers(fr, np.zeros(fr.F, np.uint8))
    assert np.array_equal(x0, fr.x) and np.array_equal(p0, fr.P)


def test_rescue_test_flags_outliers():
    b = se.make_ekf_frames(1, 32, n_id=30, outlier_ratio=0.3)
    fr = se.frame(b, 0)
    li = (~fr.outlier).astype(np.uint8); li[::3] = 0
    x, p = rne.ekf_update_inliers(fr, li)
    hi = rne.rescue_hi_inliers(fr, p, li)
    tested = (fr.ic == 1) & (li == 0)
    assert ((hi >= 0) == tested).all()               # others untouched (rescue_hi_inliers.m:37)
    assert (hi[tested & fr.outlier] == 0).all()      # gross outliers (10-40 px) never pass chi2(2, 95 %) = 5.9915


def _dense(Hc, Hf, ty, ps, n):
    H = np.zeros((2, n))
    H[:, :13] = Hc.T
    nf = 6 if ty == 0 else 3
    H[:, ps:ps + nf] = Hf[:nf].T
    return H


def test_reprediction_oracle_vs_numpy_and_finite_differences(orc):
    """rescue_hi_inliers.m:32-33 (predict_camera_measurements + calculate_derivatives at x_k_k): the C oracle against
    the independent numpy / LAPACK restatement (1e-9) and the Jacobians against central differences of the prediction
    itself (the reference's own 'Verification, OK' claim, calculate_Hi_inverse_depth_my_version.m:61-66)."""
    b = se.make_ekf_frames(1, 41, n_id=25, n_euc=9, interleave=True)
    fr = se.frame(b, 0)
    x = fr.x.copy()
    x[3:7] /= np.linalg.norm(x[3:7])                       # x_k_k leaves update.m with a unit quaternion (:44-46)
    rng = np.random.default_rng(4)
    has = np.ones(fr.F, bool); has[[3, 11]] = False        # features without a previous prediction
    h_prev = fr.h + rng.normal(size=fr.h.shape)
    # push a few features out of view: they keep their previous h (predict_camera_measurements.m:37-39)
    far = [5, 11, 20]
    for i in far:
        if fr.type[i] == 0:
            x[fr.pos[i] + 3] += 1.4                         # azimuth: beyond the 60 degree test
        else:
            x[fr.pos[i]] += 50.0
    h, has_o, pred, Hc, Hf = orc.ekf_predict(x, se.CAM, 144, 176, fr.type, fr.pos, has, h_prev)
    h2, has2, pred2, H2 = rne.predict_and_derivatives(x, se.CAM, 144, 176, fr.type, fr.pos, has, h_prev)
    assert (pred == pred2).all() and (has_o == has2).all()
    assert not pred[far].any() and pred.sum() >= fr.F - 6
    assert not has_o[11] and has_o[3] == pred[3]
    np.testing.assert_allclose(h[has_o], h2[has2], rtol=0, atol=1e-10)
    np.testing.assert_array_equal(h[~pred & has_o], h_prev[~pred & has_o])
    for i in range(fr.F):
        Hd = _dense(Hc[i], Hf[i], fr.type[i], fr.pos[i], fr.n)
        assert np.abs(Hd - H2[i]).max() <= 1e-9 * max(1.0, np.abs(H2[i]).max())
        if pred[i]:
            # the reference evaluates the distortion Jacobian at the DISTORTED pixel (dhd_dhu( camera, zi ),
            # jacob_undistor_fm_my_version.m:38): exact only without distortion (checked below), a few per cent off near the image corners with it
            num = rne.numeric_H(x, se.CAM, int(fr.type[i]), int(fr.pos[i]))
            assert np.abs(Hd - num).max() <= 0.1 * max(1.0, np.abs(num).max()), i
        if not has_o[i]:
            assert not Hd.any()
    cam0 = dict(se.CAM, k1=0.0, k2=0.0)
    h0, has0, pred0, Hc0, Hf0 = orc.ekf_predict(x, cam0, 144, 176, fr.type, fr.pos, has, h_prev)
    for i in np.flatnonzero(pred0):
        num = rne.numeric_H(x, cam0, int(fr.type[i]), int(fr.pos[i]))
        Hd = _dense(Hc0[i], Hf0[i], fr.type[i], fr.pos[i], fr.n)
        assert np.abs(Hd - num).max() <= 2e-6 * max(1.0, np.abs(num).max()), i

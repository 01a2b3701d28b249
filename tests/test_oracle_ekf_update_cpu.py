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

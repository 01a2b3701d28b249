"""Oracle of the RANSAC-pose covariance (oracle/pre3_oracle_cov.c, restating M/cov_est_RANSAC_deriv.m) against an
independent vectorised numpy restatement of the same nested central differences and against analytic facts about
E = sum |p_a - (q2R(q) p_b + T)|^2  (d2E/dT2 = 2 k I; d2E/dT dq from q2R's analytic Jacobian).  CPU only.

Tolerances: eps = 1e-6 differences amplify the rounding of E (~1e-18) by 1e12, so two correct implementations agree to
~1e-6 of the largest entry of G2tot; the covariance goes through G2tot \\ (.) and is compared at 1e-4 relative."""
import numpy as np
import pytest

EPS = 1e-6
XIDX = [10, 11, 12, 6, 7, 8, 9]


def _scene(seed, n, noise=0.004):
    rng = np.random.default_rng(seed)
    ang = rng.uniform(-0.2, 0.2, 3)
    cx, cy, cz = np.cos(ang); sx, sy, sz = np.sin(ang)
    R = (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
         @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]))
    T = rng.uniform(-0.1, 0.1, 3)
    Yb = np.column_stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1.0, 1.0, n), rng.uniform(0.8, 4.0, n)])
    Ya = Yb @ R.T + T + noise * rng.standard_normal((n, 3))
    return Ya, Yb, R, T


def _E(p):  # p (n, 13)
    bt, bp, br, at, ap, ar = (p[:, i] for i in range(6))
    a, b, c, d = (p[:, i] for i in range(6, 10))
    xa, ya, za = ar * np.cos(ap) * np.cos(at), ar * np.cos(ap) * np.sin(at), ar * np.sin(ap)
    xb, yb, zb = br * np.cos(bp) * np.cos(bt), br * np.cos(bp) * np.sin(bt), br * np.sin(bp)
    v1 = xa - ((a * a + b * b - c * c - d * d) * xb + (2 * b * c - 2 * a * d) * yb + (2 * b * d + 2 * a * c) * zb + p[:, 10])
    v2 = ya - ((2 * b * c + 2 * a * d) * xb + (a * a - b * b + c * c - d * d) * yb + (2 * c * d - 2 * a * b) * zb + p[:, 11])
    v3 = za - ((2 * b * d - 2 * a * c) * xb + (2 * c * d + 2 * a * b) * yb + (a * a - b * b - c * c + d * d) * zb + p[:, 12])
    return v1 * v1 + v2 * v2 + v3 * v3


def _grad(p):
    g = []
    for k in XIDX:
        hi, lo = p.copy(), p.copy()
        hi[:, k] = p[:, k] + EPS / 2
        lo[:, k] = p[:, k] - EPS / 2
        g.append((_E(hi) - _E(lo)) / EPS)
    return np.stack(g, -1)


def numpy_cov(orc, Ya, Yb, R, T):
    q = orc.R2q(R)
    c2s = lambda P: np.column_stack([np.arctan2(P[:, 1], P[:, 0]), np.arctan2(P[:, 2], np.hypot(P[:, 0], P[:, 1])),
                                     np.hypot(np.hypot(P[:, 0], P[:, 1]), P[:, 2])])
    n = len(Ya)
    p = np.column_stack([c2s(Yb), c2s(Ya), np.tile(q, (n, 1)), np.tile(T, (n, 1))])
    cols = []
    for k in XIDX + list(range(6)):
        hi, lo = p.copy(), p.copy()
        hi[:, k] = p[:, k] + EPS / 2
        lo[:, k] = p[:, k] - EPS / 2
        cols.append(((_grad(hi) - _grad(lo)) / EPS).sum(0))
    G2 = np.stack(cols[:7], 1)
    D = np.stack(cols[7:], 1)
    dA = np.linalg.solve(G2, D)
    sg = np.array([0.02 * np.pi / 180, 0.02 * np.pi / 180, 0.015] * 2) ** 2
    return {"G2tot": G2, "dA_dz": dA, "cov": dA @ np.diag(sg) @ dA.T, "Etot": _E(p).sum(), "Gtot": _grad(p).sum(0)}


@pytest.mark.parametrize("seed,n", [(1, 40), (2, 300), (3, 7)])
def test_oracle_vs_numpy_restatement(orc, seed, n):
    Ya, Yb, R, T = _scene(seed, n)
    o = orc.cov_est_ransac_deriv(Ya, Yb, R, T)
    r = numpy_cov(orc, Ya, Yb, R, T)
    sc = np.abs(r["G2tot"]).max()
    np.testing.assert_allclose(o["G2tot"], r["G2tot"], rtol=0, atol=2e-6 * sc)
    np.testing.assert_allclose(o["Etot"], r["Etot"], rtol=1e-12)
    np.testing.assert_allclose(o["Gtot"], r["Gtot"], rtol=0, atol=1e-8 * n)
    np.testing.assert_allclose(o["dA_dz"], r["dA_dz"], rtol=0, atol=1e-4 * np.abs(r["dA_dz"]).max())
    np.testing.assert_allclose(o["cov"], r["cov"], rtol=0, atol=1e-4 * np.abs(r["cov"]).max())
    assert o["status"] == 0 and abs(o["s2"] - o["Etot"] / (n - 3)) < 1e-15


def test_analytic_facts(orc):
    Ya, Yb, R, T = _scene(5, 200)
    o = orc.cov_est_ransac_deriv(Ya, Yb, R, T)
    n = 200
    np.testing.assert_allclose(o["G2tot"][:3, :3], 2.0 * n * np.eye(3), atol=2e-3)       # d2E/dT2 = 2 k I
    np.testing.assert_allclose(o["G2tot"], o["G2tot"].T, atol=2e-5 * np.abs(o["G2tot"]).max())  # a Hessian
    # d2E/dT dq = 2 sum d(R p_b)/dq: compare with q2R's analytic Jacobian (q2R.m:40-50)
    a, b, c, d = orc.R2q(R)
    Rq = 2 * np.array([[a, b, -c, -d], [d, c, b, a], [-c, d, -a, b], [-d, c, b, -a], [a, -b, c, -d], [b, a, d, c],
                       [c, d, a, b], [-b, -a, d, c], [a, -b, -c, d]])   # rows: R(:) column-major
    J = np.zeros((3, 4))
    for pb in Yb:
        J += 2.0 * np.stack([Rq[[0, 3, 6]].T @ pb, Rq[[1, 4, 7]].T @ pb, Rq[[2, 5, 8]].T @ pb])
    np.testing.assert_allclose(o["G2tot"][:3, 3:], J, atol=2e-5 * np.abs(J).max())
    cov = o["cov"]
    np.testing.assert_allclose(cov, cov.T, atol=1e-12 * np.abs(cov).max())
    assert (np.linalg.eigvalsh((cov + cov.T) / 2) > -1e-9 * np.abs(cov).max()).all()   # positive semi-definite
    # exact correspondences at the true pose: E = 0 and the gradient vanishes
    o0 = orc.cov_est_ransac_deriv(Yb @ R.T + T, Yb, R, T)
    assert o0["Etot"] < 1e-25 and np.abs(o0["Gtot"]).max() < 1e-6

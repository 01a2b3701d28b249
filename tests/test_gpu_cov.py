"""GPU parity of the RANSAC-pose covariance (csrc/cov.cu vs oracle/pre3_oracle_cov.c; M/cov_est_RANSAC_deriv.m), -m gpu.

Tolerance, stated: nested central differences with eps = 1e-6 amplify the rounding of E by 1e12, and CUDA's sincos /
atan2 differ from glibc's by an ulp, so G2tot agrees to 5e-6 of its largest entry and the covariance (through
G2tot \\ .) to 1e-4 of its largest entry -- the same bar the oracle meets against its independent numpy restatement."""
import importlib

import numpy as np
import pytest

from test_oracle_cov_cpu import _scene

pytestmark = pytest.mark.gpu


def _check(g, o):
    np.testing.assert_allclose(g["G2tot"], o["G2tot"], rtol=0, atol=5e-6 * np.abs(o["G2tot"]).max())
    np.testing.assert_allclose(g["Gtot"], o["Gtot"], rtol=0, atol=2e-8 * max(o["n"], 1) + 1e-9)
    np.testing.assert_allclose(g["Etot"], o["Etot"], rtol=1e-10)
    np.testing.assert_allclose(g["dA_dz"], o["dA_dz"], rtol=0, atol=1e-4 * np.abs(o["dA_dz"]).max())
    np.testing.assert_allclose(g["cov"], o["cov"], rtol=0, atol=1e-4 * np.abs(o["cov"]).max())
    np.testing.assert_allclose(g["s2"], o["s2"], rtol=1e-10)
    assert g["n"] == o["n"] and g["status"] == 0


def test_cov_batch_vs_oracle(ctx, orc):
    P, Nmax = 5, 400
    ns = [400, 37, 4, 250, 399]
    Ya = np.zeros((P, Nmax, 3)); Yb = np.zeros((P, Nmax, 3)); R = np.zeros((P, 3, 3)); T = np.zeros((P, 3))
    for p in range(P):
        a, b, R[p], T[p] = _scene(20 + p, ns[p])
        Ya[p, : ns[p]], Yb[p, : ns[p]] = a, b
    out = ctx.cov_est_ransac_batch(Ya, Yb, R, T, n_corr=np.array(ns, np.int32))
    for p in range(P):
        _check(out[p], orc.cov_est_ransac_deriv(Ya[p, : ns[p]], Yb[p, : ns[p]], R[p], T[p]))


def test_cov_support_set_mask_and_matlab_mirror(ctx, orc):
    M = importlib.import_module("3pre_b200.matlab")
    Ya, Yb, R, T = _scene(31, 300)
    rng = np.random.default_rng(3)
    mask = rng.uniform(size=300) < 0.6
    Ya[~mask] += 0.5      # outliers: not in the support set
    out = ctx.cov_est_ransac_batch(Ya[None], Yb[None], R[None], T[None], masks=mask[None])[0]
    o = orc.cov_est_ransac_deriv(Ya[mask], Yb[mask], R, T)
    _check(out, o)
    res = M.cov_est_RANSAC_deriv(Ya[mask].T, Yb[mask].T, R, T)
    np.testing.assert_allclose(res["sm_cov_censi"], o["cov"], rtol=0, atol=1e-4 * np.abs(o["cov"]).max())
    # empty support set
    e = ctx.cov_est_ransac_batch(Ya[None], Yb[None], R[None], T[None], masks=np.zeros((1, 300), bool))[0]
    assert e["status"] == 1 and e["n"] == 0


def test_cov_on_the_records_of_a_ransac_batch(ctx, orc, pre3):
    """The _dev form run directly on the records and masks pre3_ransac_batch_dev left on the device (N = 20 000 in one
    pair: the point loop is split over many blocks)."""
    import torch
    N = 20000
    rng = np.random.default_rng(77)
    Ya, Yb, _, _ = _scene(77, N, noise=0.003)
    bad = rng.uniform(size=N) < 0.5
    Ya[bad] += rng.uniform(-1, 1, (int(bad.sum()), 3))
    dYa, dYb = torch.from_numpy(Ya[None].copy()).cuda(), torch.from_numpy(Yb[None].copy()).cuda()
    opts = pre3.make_opts(H=300, seed=4, adaptive=False, distance_threshold=0.02)
    res = torch.zeros(1, 240, dtype=torch.uint8, device="cuda")
    masks = torch.zeros(1, N, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.ransac_batch_dev(dYa, dYb, opts, res, masks=masks)
    out = torch.zeros(1, pre3.COV_RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    rt = res.view(torch.float64)[:, 6:]          # R starts at byte 48 of pre3_pair_result
    ctx.cov_est_ransac_batch_dev(dYa, dYb, rt, 30, out, masks=masks)
    ctx.sync()
    rec = np.frombuffer(res.cpu().numpy().tobytes(), dtype=pre3.RESULT_DTYPE)[0]
    g = pre3.unpack_cov(np.frombuffer(out.cpu().numpy().tobytes(), dtype=pre3.COV_RESULT_DTYPE)[0])
    mk = masks[0].cpu().numpy().astype(bool)
    assert rec["status"] == 0 and mk.sum() > 5000
    o = orc.cov_est_ransac_deriv(Ya[mk], Yb[mk], rec["R"].reshape(3, 3).T, rec["T"])
    _check(g, o)

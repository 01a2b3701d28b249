"""GPU parity: the EKF partial updates around ransac_hypotheses (first part) through the C ABI vs
the dense numpy / LAPACK restatement of M/update.m, ekf_update_li_inliers.m, ekf_update_hi_inliers.m and the test of
rescue_hi_inliers.m (oracle/ref_numpy_ekf.py).  Floating-point path: tolerance 1e-9 relative to max|P| (stated in
DESIGN.md); the flags of the rescue test must agree wherever the statistic is not within 1e-6 of the 5.9915 bound."""
import importlib

import numpy as np
import pytest

from oracle import ref_numpy_ekf as rne

pytestmark = pytest.mark.gpu
se = importlib.import_module("3pre_b200.synth_ekf")
TOL = 1e-9


def _check_update(fr, flags, xo, Po, m, x=None, P=None):
    gx, gP = rne.ekf_update_inliers(fr, flags, x, P)
    assert m == 2 * int((flags == 1).sum())
    scale = np.abs(gP).max()
    assert np.abs(Po.T - gP).max() <= TOL * scale, np.abs(Po.T - gP).max() / scale
    assert np.abs(xo - gx).max() <= TOL * max(1.0, np.abs(gx).max())
    return gx, gP


@pytest.mark.parametrize("n_id,n_euc", [(20, 0), (12, 9), (70, 5)])
def test_li_update_vs_numpy(ctx, n_id, n_euc):
    Fr = 4
    b = se.make_ekf_frames(Fr, 600 + n_id, n_id=n_id, n_euc=n_euc, interleave=True)
    fb = se.batch_to_numpy(b)
    flags = (~fb["outlier"]).astype(np.uint8)
    flags[1, ::3] = 0          # ragged: different m per frame
    flags[2] = 0               # nothing flagged: the frame is copied through (update.m:50-54)
    flags[3, 5] = 2            # only == 1 counts (ekf_update_li_inliers.m:17)
    xo, Po, m = ctx.ekf_update_batch(fb, flags)
    for f in range(Fr):
        fr = se.frame(b, f)
        _check_update(fr, flags[f], xo[f], Po[f], m[f])
        if f == 2:
            assert np.array_equal(xo[f], fb["x"][f]) and np.array_equal(Po[f], fb["P"][f])
        else:
            assert abs(np.linalg.norm(xo[f, 3:7]) - 1) < 1e-14          # q / |q| (update.m:48)
            assert np.abs(Po[f] - Po[f].T).max() <= 1e-12 * np.abs(Po[f]).max()
            assert np.trace(Po[f]) < np.trace(fb["P"][f])                # information was added


def test_li_then_rescue_then_hi_update(ctx):
    """The sequence of mono_slam.m around ransac_hypotheses: li update -> rescue test -> hi update, with the
    measurements NOT re-predicted in between (that part of rescue_hi_inliers.m:32-33 stays with the caller)."""
    torch = pytest.importorskip("torch")
    Fr = 3
    b = se.make_ekf_frames(Fr, 777, device="cuda", n_id=40, outlier_ratio=0.3)
    b["cam"] = dict(se.CAM)
    F, n = b["F"], b["n"]
    li = (~b["outlier"]).to(torch.uint8)
    li[:, ::4] = 0                                    # some inliers are left for the rescue step
    x1 = torch.zeros_like(b["x"]); P1 = torch.zeros_like(b["P"])
    ctx.ekf_update_batch_dev(b, li, x1, P1)
    hi = torch.full((Fr, F), 7, dtype=torch.uint8, device="cuda")
    ctx.ekf_rescue_hi_inliers_batch_dev(b, P1, li, hi)
    ctx.sync()                                        # the context has its own stream: torch must see its results
    x2 = torch.zeros_like(x1); P2 = torch.zeros_like(P1)
    sel2 = (hi == 1).to(torch.uint8)
    torch.cuda.synchronize()
    ctx.ekf_update_batch_dev(b, sel2, x2, P2, x=x1, P=P1)
    ctx.sync()
    for f in range(Fr):
        fr = se.frame(b, f)
        lf = li[f].cpu().numpy()
        gx, gP = _check_update(fr, lf, x1[f].cpu().numpy(), P1[f].cpu().numpy(), 2 * int(lf.sum()))
        g_hi = rne.rescue_hi_inliers(fr, gP, lf)
        got = hi[f].cpu().numpy().astype(np.int32)
        got[got == 7] = -1                             # untouched entries
        # statistic of every tested feature, to exclude razor-edge cases from the flag comparison
        for i in np.flatnonzero(g_hi >= 0):
            Hi = rne.dense_H(fr, i)
            nu = fr.z[i] - fr.h[i]
            q = nu @ np.linalg.inv(Hi @ gP @ Hi.T) @ nu
            if abs(q - 5.9915) > 1e-6:
                assert got[i] == g_hi[i]
        np.testing.assert_array_equal(got < 0, g_hi < 0)
        s2 = sel2[f].cpu().numpy()
        _check_update(fr, s2, x2[f].cpu().numpy(), P2[f].cpu().numpy(), 2 * int(s2.sum()), gx, gP)


def test_full_size_frame(ctx):
    """BASELINE config 4 shape: 200 inverse-depth features, n = 1213."""
    b = se.make_ekf_frames(2, 4242, n_id=200)
    fb = se.batch_to_numpy(b)
    flags = (~fb["outlier"]).astype(np.uint8)
    xo, Po, m = ctx.ekf_update_batch(fb, flags)
    for f in range(2):
        _check_update(se.frame(b, f), flags[f], xo[f], Po[f], m[f])
    assert m.min() > 250


def test_chunked_batches_and_degenerate_sizes(pre3, monkeypatch):
    """Batches larger than the workspace budget run in chunks (forced here: 2 frames per chunk); no features at all."""
    monkeypatch.setenv("PRE3_UPD_CHUNK", "2")
    Fr = 5
    b = se.make_ekf_frames(Fr, 91, n_id=10, n_euc=4)
    fb = se.batch_to_numpy(b)
    flags = (~fb["outlier"]).astype(np.uint8)
    with pre3.Context(0) as c2:
        xo, Po, m = c2.ekf_update_batch(fb, flags)
        for f in range(Fr):
            _check_update(se.frame(b, f), flags[f], xo[f], Po[f], m[f])
        empty = {k: (v[:, :0] if isinstance(v, np.ndarray) and v.ndim >= 2 and k not in ("x", "P") else v)
                 for k, v in fb.items()}
        xo, Po, m = c2.ekf_update_batch(empty, np.zeros((Fr, 0), np.uint8))
        assert (m == 0).all() and np.array_equal(xo, fb["x"]) and np.array_equal(Po, fb["P"])


def test_argument_errors(ctx, pre3):
    b = se.batch_to_numpy(se.make_ekf_frames(1, 1, n_id=4))
    with pytest.raises(pre3.Pre3Error):
        ctx.ekf_update_batch(b, np.zeros((1, 4), np.uint8), x=np.zeros((1, 5)), P=np.zeros((1, 5, 5)))

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


def test_update_with_the_functions_own_signature(ctx):
    """[x, p, K] = update(x, p, H, R, z, h) with dense H and a general R (M/update.m:27), through the mirror."""
    ml = importlib.import_module("3pre_b200.matlab")
    b = se.make_ekf_frames(1, 55, n_id=25, n_euc=4, interleave=True)
    fr = se.frame(b, 0)
    sel = [i for i in range(fr.F) if not fr.outlier[i]]
    H = np.vstack([rne.dense_H(fr, i) for i in sel])
    z = np.concatenate([fr.z[i] for i in sel]); h = np.concatenate([fr.h[i] for i in sel])
    rng = np.random.default_rng(1)
    A = rng.normal(size=(len(z), len(z))) * 0.1
    R = np.eye(len(z)) * 1.5 + A @ A.T                       # a general SPD measurement covariance
    gx, gP, gK = rne.update(fr.x, fr.P, H, R, z, h)
    x, P, K = ml.update(fr.x.reshape(-1, 1), fr.P, H, R, z.reshape(-1, 1), h.reshape(-1, 1))
    assert x.shape == (fr.n, 1) and np.abs(x.ravel() - gx).max() < TOL
    assert np.abs(P - gP).max() <= TOL * np.abs(gP).max() and np.abs(K - gK).max() <= TOL * np.abs(gK).max()
    from scipy import sparse
    x2, P2, _ = ml.update(fr.x, fr.P, sparse.csr_matrix(H), R, z, h)
    assert np.array_equal(x2, x.ravel()) and np.array_equal(P2, P)
    x0, P0, K0 = ml.update(fr.x, fr.P, np.zeros((0, fr.n)), np.zeros((0, 0)), np.zeros(0), np.zeros(0))
    assert np.array_equal(x0, fr.x) and np.array_equal(P0, fr.P) and K0 == 0


def test_mex_gateway_update_shadows_update_m():
    """3pre_b200/mex_files/update.cpp linked against the stub MEX runtime (oracle/mex_stub) and called through
    mexFunction with mxArrays: the drop-in for M/update.m itself."""
    import ctypes as C
    import os
    import subprocess
    from oracle import refmex
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libpre3_update_gw.so")
    libdir = os.path.join(root, "3pre_b200", "lib")
    subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-o", so, os.path.join(root, "3pre_b200", "mex_files", "update.cpp"),
                    "-x", "c", os.path.join(root, "oracle", "mex_stub", "mex_stub.c"), "-x", "none",
                    "-I", os.path.join(root, "oracle", "mex_stub"), "-I", os.path.join(root, "3pre_b200", "mex_files"),
                    "-I", os.path.join(root, "include"), "-L", libdir, "-lpre3", f"-Wl,-rpath,{libdir}"], check=True)
    L = refmex.lib(so)

    def call(args, nout):   # args: MATLAB matrices (rows, cols); passed column-major
        keep = [np.ascontiguousarray(np.atleast_2d(a).T, np.float64) for a in args]
        ins = [L.stub_wrap(6, k.shape[1], k.shape[0], k.ctypes.data) for k in keep]
        in_arr = (C.POINTER(refmex._MxArray) * len(ins))(*ins)
        out_arr = (C.POINTER(refmex._MxArray) * nout)()
        rc = L.stub_call_mex(nout, out_arr, len(ins), in_arr)
        for a in ins:
            L.mxDestroyArray(a)
        if rc != 0:
            raise refmex.MexError(L.stub_last_error().decode())
        res = []
        for i in range(nout):
            m = out_arr[i].contents
            res.append(np.ctypeslib.as_array(C.cast(m.data, C.POINTER(C.c_double)), shape=(m.n, m.m)).copy().T)
            L.mxDestroyArray(out_arr[i])
        return res

    b = se.make_ekf_frames(1, 56, n_id=12)
    fr = se.frame(b, 0)
    sel = [i for i in range(fr.F) if not fr.outlier[i]]
    H = np.vstack([rne.dense_H(fr, i) for i in sel])
    z = np.concatenate([fr.z[i] for i in sel]); h = np.concatenate([fr.h[i] for i in sel])
    x, P, K = call([fr.x.reshape(-1, 1), fr.P, H, np.eye(len(z)), z.reshape(-1, 1), h.reshape(-1, 1)], 3)
    gx, gP, gK = rne.update(fr.x, fr.P, H, np.eye(len(z)), z, h)
    assert x.shape == (fr.n, 1) and np.abs(x.ravel() - gx).max() < TOL
    assert np.abs(P - gP).max() <= TOL * np.abs(gP).max() and np.abs(K - gK).max() <= TOL * np.abs(gK).max()
    # empty measurement vector: copied through, K = 0 (update.m:50-54)
    x0, P0, K0 = call([fr.x.reshape(-1, 1), fr.P, np.zeros((0, fr.n)), np.zeros((0, 0)), np.zeros((0, 1)), np.zeros((0, 1))], 3)
    assert np.array_equal(x0.ravel(), fr.x) and np.array_equal(P0, fr.P) and K0.ravel()[0] == 0
    with pytest.raises(refmex.MexError, match="six inputs"):
        call([fr.x.reshape(-1, 1), fr.P], 1)
    with pytest.raises(refmex.MexError, match="H must be"):
        call([fr.x.reshape(-1, 1), fr.P, H[:, :5], np.eye(len(z)), z.reshape(-1, 1), h.reshape(-1, 1)], 1)


def test_reprediction_at_x_k_k_vs_oracle(ctx, orc):
    """rescue_hi_inliers.m:32-33 on the GPU (k_ekf_predict) against the C oracle sharing its specification: flags and
    h bit for bit (same sincos / operation order), H to 1e-12 of its largest entry (CUDA's atan2 only feeds the
    comparisons; -fmad=false keeps the products unfused), and against the independent numpy restatement at 1e-9."""
    Fr = 3
    b = se.make_ekf_frames(Fr, 91, n_id=60, n_euc=20, interleave=True)
    fb = se.batch_to_numpy(b)
    x = fb["x"].copy()
    x[:, 3:7] /= np.linalg.norm(x[:, 3:7], axis=1, keepdims=True)
    rng = np.random.default_rng(8)
    has = rng.uniform(size=fb["type"].shape) < 0.9
    h_prev = fb["h"] + rng.normal(size=fb["h"].shape)
    for f in range(Fr):                                   # push some features out of view
        for i in rng.choice(fb["type"].shape[1], 6, replace=False):
            if fb["type"][f, i] == 0:
                x[f, fb["pos"][f, i] + 3] += 1.4
            else:
                x[f, fb["pos"][f, i]] += 50.0
    h, has_o, pred, Hc, Hf = ctx.ekf_predict_measurements_batch(x, se.CAM, 144, 176, fb["type"], fb["pos"], has, h_prev)
    n_unpred = 0
    for f in range(Fr):
        oh, ohas, opred, oHc, oHf = orc.ekf_predict(x[f], se.CAM, 144, 176, fb["type"][f], fb["pos"][f], has[f], h_prev[f])
        np.testing.assert_array_equal(pred[f], opred)
        np.testing.assert_array_equal(has_o[f], ohas)
        np.testing.assert_array_equal(h[f][ohas], oh[ohas])
        sc = max(np.abs(oHc).max(), np.abs(oHf).max())
        np.testing.assert_allclose(Hc[f], oHc, rtol=0, atol=1e-12 * sc)
        np.testing.assert_allclose(Hf[f], oHf, rtol=0, atol=1e-12 * sc)
        n_unpred += int((~opred).sum())
        h2, has2, pred2, H2 = rne.predict_and_derivatives(x[f], se.CAM, 144, 176, fb["type"][f], fb["pos"][f], has[f], h_prev[f])
        for i in np.flatnonzero(has2):
            Hd = np.zeros((2, x.shape[1]))
            Hd[:, :13] = Hc[f, i].T
            nf = 6 if fb["type"][f, i] == 0 else 3
            Hd[:, fb["pos"][f, i]: fb["pos"][f, i] + nf] = Hf[f, i, :nf].T
            assert np.abs(Hd - H2[i]).max() <= 1e-9 * max(1.0, np.abs(H2[i]).max())
    assert n_unpred >= 10


def test_rescue_with_reprediction_device_sequence(ctx):
    """li update -> re-prediction at x_k_k -> chi2 test -> hi update, all on device buffers: the whole of
    rescue_hi_inliers.m, compared with the numpy restatements chained the same way."""
    torch = pytest.importorskip("torch")
    Fr = 2
    b = se.make_ekf_frames(Fr, 778, device="cuda", n_id=40, n_euc=8, interleave=True, outlier_ratio=0.3)
    b["cam"] = dict(se.CAM)
    F, n = b["F"], b["n"]
    li = (~b["outlier"]).to(torch.uint8)
    li[:, ::4] = 0
    x1 = torch.zeros_like(b["x"]); P1 = torch.zeros_like(b["P"])
    ctx.ekf_update_batch_dev(b, li, x1, P1)
    h1 = torch.zeros_like(b["h"]); Hc1 = torch.zeros_like(b["Hcam"]); Hf1 = torch.zeros_like(b["Hfeat"])
    has1 = torch.zeros(Fr, F, dtype=torch.uint8, device="cuda"); pr1 = torch.zeros_like(has1)
    ctx.ekf_predict_measurements_batch_dev(x1, se.CAM, 144, 176, b["type"], b["pos"], None, b["h"], h1, has1, pr1, Hc1, Hf1)
    hi = torch.full((Fr, F), 7, dtype=torch.uint8, device="cuda")
    ctx.ekf_rescue_hi_inliers_batch_dev(b, P1, li, hi, h=h1, Hcam=Hc1, Hfeat=Hf1)
    ctx.sync()
    for f in range(Fr):
        fr = se.frame(b, f)
        lf = li[f].cpu().numpy()
        gx, gP = rne.ekf_update_inliers(fr, lf)
        h2, has2, pred2, H2 = rne.predict_and_derivatives(gx, se.CAM, 144, 176, fr.type, fr.pos, np.ones(F, bool), fr.h)
        assert pred2.sum() >= F - 4
        np.testing.assert_allclose(h1[f].cpu().numpy(), h2, rtol=0, atol=1e-7)   # x1 vs gx differ by rounding of the update
        got = hi[f].cpu().numpy().astype(np.int32)
        for i in range(F):
            tested = fr.ic[i] == 1 and lf[i] == 0
            assert (got[i] != 7) == tested
            if tested:
                nu = fr.z[i] - h2[i]
                q = nu @ np.linalg.inv(H2[i] @ gP @ H2[i].T) @ nu
                if abs(q - 5.9915) > 1e-4:
                    assert got[i] == (1 if q < 5.9915 else 0)


def test_matlab_rescue_hi_inliers_mirror(ctx):
    """The MATLAB-shaped mirror of rescue_hi_inliers(filter, features_info, cam) on a features_info list."""
    M = importlib.import_module("3pre_b200.matlab")
    b = se.make_ekf_frames(1, 779, n_id=20, n_euc=6, interleave=True, outlier_ratio=0.3)
    fr = se.frame(b, 0)
    li = (~fr.outlier).astype(np.uint8); li[::3] = 0
    x1, P1 = rne.ekf_update_inliers(fr, li)
    cam = dict(se.CAM, nRows=144, nCols=176)
    feats = []
    for i in range(fr.F):
        feats.append({"type": "inversedepth" if fr.type[i] == 0 else "cartesian", "h": fr.h[i].reshape(1, 2),
                      "z": fr.z[i].copy(), "individually_compatible": int(fr.ic[i]), "low_innovation_inlier": int(li[i]),
                      "H": rne.dense_H(fr, i)})
    out = M.rescue_hi_inliers({"x_k_k": x1, "p_k_k": P1}, feats, cam)
    h2, has2, pred2, H2 = rne.predict_and_derivatives(x1, se.CAM, 144, 176, fr.type, fr.pos, np.ones(fr.F, bool), fr.h)
    n_tested = 0
    for i in range(fr.F):
        np.testing.assert_allclose(out[i]["h"].reshape(-1), h2[i], rtol=0, atol=1e-10)
        assert np.abs(out[i]["H"] - H2[i]).max() <= 1e-9 * max(1.0, np.abs(H2[i]).max())
        tested = fr.ic[i] == 1 and li[i] == 0
        assert ("high_innovation_inlier" in out[i]) == tested
        if tested:
            n_tested += 1
            nu = fr.z[i] - h2[i]
            q = nu @ np.linalg.inv(H2[i] @ P1 @ H2[i].T) @ nu
            if abs(q - 5.9915) > 1e-4:
                assert out[i]["high_innovation_inlier"] == (1 if q < 5.9915 else 0)
    assert n_tested >= 5

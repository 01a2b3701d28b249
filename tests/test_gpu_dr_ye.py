"""GPU parity: the code_from_dr_ye variant through the C ABI vs the CPU oracle
(oracle/pre3_oracle_dr_ye.c) and the committed golden vectors of the independent numpy restatement.
Bit-exact: tmp_cnum of every iteration, the winner, op_num, the support set, nIterationRansac, the seeded
sample sets; 1e-9: refit rotation / translation, ErrorMean, ErrorStd.  `M/` = /root/reference/matlab_code/.
"""
import importlib
import os

import numpy as np
import pytest

from oracle import ref_numpy as rn

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "dr_ye.npz"))


def _check(o, rec, mask, st, counts, N, tol=1e-9):
    assert rec["status"] == o.status and rec["n_matches"] == N
    if counts is not None:
        np.testing.assert_array_equal(counts, o.counts)
    assert rec["best_fit"] == o.op_num and rec["n_consumed"] == o.n_loops
    assert rec["n_iter"] == o.n_iteration_ransac == st["n_iteration_ransac"] and st["n_loops"] == o.n_loops
    if o.status in (1, 5):
        return
    assert rec["best_sample"] == o.best_sample and rec["thr"] == o.thr
    if o.status != 0:
        assert not mask.any()
        return
    np.testing.assert_array_equal(mask[:N].astype(bool), o.mask)
    assert not mask[N:].any()
    assert rec["state"] == o.state
    R = np.array(rec["R"]).reshape(3, 3).T
    Rh = np.array(rec["R_hyp"]).reshape(3, 3).T
    np.testing.assert_array_equal(Rh, o.R_hyp)            # minimal fit: same operation order, same bits
    np.testing.assert_array_equal(np.array(rec["T_hyp"]), o.T_hyp)
    assert rn.rot_angle(R, o.R) < tol and np.abs(np.array(rec["T"]) - o.T).max() < tol
    assert abs(st["error_mean"] - o.error_mean) < tol and abs(st["error_std"] - o.error_std) < tol
    assert abs(rec["error_sum"] - o.error_sum) < tol * max(1, o.op_num)
    assert abs(st["dist"] * 0.001 - o.thr) < 1e-15


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_golden(ctx, gold, name):
    Ya, Yb, draws = gold[f"{name}_Ya"], gold[f"{name}_Yb"], gold[f"{name}_draws"]
    res, masks, st, counts = ctx.vodometry_dr_ye_batch(Ya[None], Yb[None], samples=draws[None], want_counts=True)
    status, op_num, best, n_loops, nit, state = gold[f"{name}_scalars"]
    r = res[0]
    assert (r["status"], r["best_fit"], r["best_sample"], r["n_consumed"], r["n_iter"], r["state"]) == \
        (status, op_num, best, n_loops, nit, state)
    np.testing.assert_array_equal(counts[0], gold[f"{name}_counts"])
    np.testing.assert_array_equal(masks[0].astype(bool), gold[f"{name}_mask"])
    thr, mean, std = gold[f"{name}_stats"]
    assert abs(r["thr"] - thr) < 1e-15 and abs(st[0]["error_mean"] - mean) < 1e-9 and abs(st[0]["error_std"] - std) < 1e-9
    assert rn.rot_angle(np.array(r["R"]).reshape(3, 3).T, gold[f"{name}_R"]) < 1e-9
    assert np.abs(np.array(r["T"]) - gold[f"{name}_T"]).max() < 1e-9


def test_batch_vs_oracle_supplied_draws(ctx, orc, synth):
    P, Nmax, H = 12, 320, 700
    Ya = np.zeros((P, Nmax, 3))
    Yb = np.zeros((P, Nmax, 3))
    n = np.zeros(P, np.int32)
    draws = np.zeros((P, H, 4), np.int32)
    sizes = [320, 300, 150, 64, 13, 12, 7, 4, 3, 0, 200, 90]   # ragged, incl. C(n,4) < 700 and pnum < 4
    for p in range(P):
        N = sizes[p]
        n[p] = N
        if N:
            c = synth.make_correspondences(4000 + p, N=N, outlier_ratio=[0.3, 0.5, 0.6, 0.1][p % 4])
            Ya[p, :N], Yb[p, :N] = c.Ya, c.Yb
            draws[p] = synth.make_draws(4100 + p, H, max(N, 4)) % max(N, 1)
    Yb[10, :200] *= 0.05   # nothing farther than 0.4 m -> status 5
    Ya[10, :200] *= 0.05
    res, masks, st, counts = ctx.vodometry_dr_ye_batch(Ya, Yb, n_corr=n, samples=draws, want_counts=True)
    seen = set()
    for p in range(P):
        N = sizes[p]
        o = orc.vodometry_dr_ye(Ya[p, :N], Yb[p, :N], samples=draws[p])
        _check(o, res[p], masks[p], st[p], counts[p], N)
        seen.add(o.status)
    assert {0, 1, 5} <= seen


def test_seeded_sampler_and_match_ids(ctx, orc, synth):
    # match ids with repeated k2 and with k1 values that also occur as k2 values (the mixed-row tests of :35)
    P, N, H = 6, 180, 700
    rng = np.random.default_rng(77)
    Ya = np.zeros((P, N, 3))
    Yb = np.zeros((P, N, 3))
    match = np.zeros((P, N, 2), np.int32)
    for p in range(P):
        c = synth.make_correspondences(4300 + p, N=N, outlier_ratio=0.4)
        Ya[p], Yb[p] = c.Ya, c.Yb
        match[p, :, 0] = np.sort(rng.choice(400, N, replace=False))
        match[p, :, 1] = rng.integers(0, 60 if p % 2 else 400, N)
    res, masks, st, counts = ctx.vodometry_dr_ye_batch(Ya, Yb, match=match, seed=99, want_counts=True)
    for p in range(P):
        o = orc.vodometry_dr_ye(Ya[p], Yb[p], match=match[p], seed=99, pair=p)
        _check(o, res[p], masks[p], st[p], counts[p], N)
    # no match ids: match(:,i) = [i;i]
    res, masks, st, counts = ctx.vodometry_dr_ye_batch(Ya[:2], Yb[:2], seed=5, want_counts=True)
    for p in range(2):
        _check(orc.vodometry_dr_ye(Ya[p], Yb[p], seed=5, pair=p), res[p], masks[p], st[p], counts[p], N)


def test_failed_fits_are_scored(ctx, orc, synth):
    c = synth.make_correspondences(21, N=60, outlier_ratio=0.2)
    Yb = c.Yb.copy()
    Yb[:4] = np.outer(np.arange(4.0), [1.0, 1.0, 1.0]) + [0, 0, 2]   # collinear -> state -1, rot = H, trans = 0
    Ya = Yb @ c.R.T + c.t
    draws = synth.make_draws(22, 64, 60)
    draws[0] = [0, 1, 2, 3]
    draws[1] = [3, 3, 3, 3]                                            # all the same point: H = 0
    res, masks, st, counts = ctx.vodometry_dr_ye_batch(Ya[None], Yb[None], samples=draws[None], want_counts=True)
    o = orc.vodometry_dr_ye(Ya, Yb, samples=draws)
    assert o.counts[0] >= 0 and o.counts[1] >= 0
    _check(o, res[0], masks[0], st[0], counts[0], 60)


def test_no_consensus(ctx, orc, synth):
    rng = np.random.default_rng(3)
    A = rng.normal(size=(1, 6, 3)) * 5 + [0, 0, 8]
    B = rng.normal(size=(1, 6, 3)) * 5 + [0, 0, 8]
    draws = (synth.make_draws(5, 15, 6))[None]
    res, masks, st, counts = ctx.vodometry_dr_ye_batch(A, B, samples=draws, want_counts=True)
    o = orc.vodometry_dr_ye(A[0], B[0], samples=draws[0])
    _check(o, res[0], masks[0], st[0], counts[0], 6)


def test_pairs_pipeline_with_the_dr_ye_variant(ctx, pre3, orc, synth):
    # siftmatch -> gather -> dr_ye RANSAC through pre3_pairs (opts.method = DR_YE): match ids = the siftmatch
    # output, seeded sampler; compared with the oracle run stage by stage on the same pair
    L = importlib.import_module("3pre_b200._lib")
    P = 5
    fps = [synth.make_frame_pair(5200 + p, K1=256, K2=256, n_corr=150, outlier_ratio=0.3) for p in range(P)]
    d1 = np.stack([f.desc1 for f in fps]); d2 = np.stack([f.desc2 for f in fps])
    x1 = np.stack([f.xyz1 for f in fps]); x2 = np.stack([f.xyz2 for f in fps])
    o = pre3.make_opts(method=L.METHOD_DR_YE, k=4, max_iteration=700, H=700, seed=31)
    res, matches, masks = ctx.pairs(d1, d2, x1, x2, o, pair_id0=40)
    for p in range(P):
        pairs, _ = orc.siftmatch(d1[p], d2[p], 1.5)
        n = pairs.shape[0]
        assert res[p]["n_matches"] == n
        np.testing.assert_array_equal(matches[p, :n], pairs)
        g = orc.vodometry_dr_ye(x1[p][pairs[:, 0]], x2[p][pairs[:, 1]], match=pairs, seed=31, pair=40 + p)
        st = {"n_iteration_ransac": g.n_iteration_ransac, "n_loops": g.n_loops, "error_mean": g.error_mean,
              "error_std": g.error_std, "dist": g.thr / 0.001}
        _check(g, res[p], masks[p], st, None, n)
        assert g.status == 0 and rn.rot_angle(np.array(res[p]["R"]).reshape(3, 3).T, fps[p].R) < 5e-3


def test_matlab_mirror_vodometry_dr_ye(orc, synth):
    # frames with depth maps: pset = [-x(ROW,COL); -y(ROW,COL); z(ROW,COL)] at the rounded feature positions
    ml = importlib.import_module("3pre_b200.matlab")
    rng = np.random.default_rng(8)
    f = synth.make_frame_pair(6100, K1=200, K2=200, n_corr=120, outlier_ratio=0.25)

    def frame(desc, xyz):
        K = desc.shape[0]
        cells = rng.permutation(144 * 176)[:K]
        row, col = cells // 176, cells % 176
        x = rng.normal(size=(144, 176)); y = rng.normal(size=(144, 176)); z = rng.uniform(1, 5, size=(144, 176))
        x[row, col], y[row, col], z[row, col] = -xyz[:, 0], -xyz[:, 1], xyz[:, 2]
        frm = np.stack([col + rng.uniform(-0.45, 0.45, K), row + rng.uniform(-0.45, 0.45, K), np.ones(K), np.zeros(K)])
        cm = np.full((144, 176), 100.0)
        return {"frm": frm, "des": desc.T.copy(), "x": x, "y": y, "z": z, "confidence_map": cm}

    D1, D2 = frame(f.desc1, f.xyz1), frame(f.desc2, f.xyz2)
    D2["confidence_map"][0, 0] = 1000.0   # every feature of frame 2 is now below half the maximum ...
    out = ml.vodometry_dr_ye(D1, D2, confidence_map=True)
    assert out[5] == 1 and out[11]["SolutionState"] == 4 and out[11]["nF2_Confidence_Filtered"] == 0   # ... :152-160
    rot, phi, theta, psi, trans, err, pnum, op_num, sta, p1, p2, stat = ml.vodometry_dr_ye(D1, D2, seed=3)
    pairs, _ = orc.siftmatch(f.desc1, f.desc2, 1.5)
    g = orc.vodometry_dr_ye(f.xyz1[pairs[:, 0]], f.xyz2[pairs[:, 1]], match=pairs, seed=3, pair=0)
    assert err == 0 and sta == 1 and pnum == pairs.shape[0] and op_num == g.op_num == p1.shape[1] == p2.shape[1]
    assert rn.rot_angle(rot, g.R) < 1e-9 and np.abs(trans.ravel() - g.T).max() < 1e-9
    np.testing.assert_allclose(p1.T, f.xyz1[pairs[:, 0]][g.mask], atol=0)
    assert abs(stat["ErrorMean"] - g.error_mean) < 1e-9 and abs(stat["ErrorStd"] - g.error_std) < 1e-9
    assert stat["nIterationRansac"] == g.n_iteration_ransac and abs(stat["InlierRatio"] - 100 * op_num / pnum) < 1e-12
    np.testing.assert_allclose([phi, theta, psi], ml.R2e(rot))
    T, q, R, sta2, _ = ml.Calculate_V_Omega_RANSAC_dr_ye(D1, D2, seed=3)
    assert sta2 == 1 and np.array_equal(R, rot) and np.array_equal(T, trans) and abs(np.linalg.norm(q) - 1) < 1e-12


def test_device_entry_is_stream_ordered(ctx, pre3, orc, synth):
    torch = pytest.importorskip("torch")
    L = importlib.import_module("3pre_b200._lib")
    P, N = 70, 128    # P >= 64
    cs = [synth.make_correspondences(4700 + p, N=N, outlier_ratio=0.35) for p in range(P)]
    Ya = torch.tensor(np.stack([c.Ya for c in cs]), device="cuda")
    Yb = torch.tensor(np.stack([c.Yb for c in cs]), device="cuda")
    res = torch.zeros(P, 240, dtype=torch.uint8, device="cuda")
    masks = torch.zeros(P, N, dtype=torch.uint8, device="cuda")
    stat = torch.zeros(P, 32, dtype=torch.uint8, device="cuda")
    o = pre3.make_opts(method=L.METHOD_DR_YE, k=4, max_iteration=700, H=700, seed=2)
    ctx.vodometry_dr_ye_batch_dev(Ya, Yb, o, res, masks=masks, stat=stat, pair_id0=1000)
    ctx.sync()
    rec = np.frombuffer(res.cpu().numpy().tobytes(), dtype=pre3.RESULT_DTYPE)
    st = np.frombuffer(stat.cpu().numpy().tobytes(), dtype=pre3.DR_YE_STAT_DTYPE)
    mk = masks.cpu().numpy()
    for p in range(0, P, 7):
        g = orc.vodometry_dr_ye(cs[p].Ya, cs[p].Yb, seed=2, pair=1000 + p)
        _check(g, rec[p], mk[p], st[p], None, N)


def test_mex_gateway_vodometry_dr_ye(orc, synth):
    """3pre_b200/mex_files/vodometry_dr_ye_mex.cpp linked against the stub MEX runtime that also drives the
    reference's own siftmatch.c gateway (oracle/mex_stub): called through mexFunction with mxArrays."""
    import ctypes as C
    import subprocess
    from oracle import refmex
    root = ROOT
    out = os.path.join(root, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libpre3_vodometry_dr_ye_gw.so")
    libdir = os.path.join(root, "3pre_b200", "lib")
    subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-o", so,
                    os.path.join(root, "3pre_b200", "mex_files", "vodometry_dr_ye_mex.cpp"),
                    "-x", "c", os.path.join(root, "oracle", "mex_stub", "mex_stub.c"), "-x", "none",
                    "-I", os.path.join(root, "oracle", "mex_stub"), "-I", os.path.join(root, "3pre_b200", "mex_files"),
                    "-I", os.path.join(root, "include"), "-L", libdir, "-lpre3", f"-Wl,-rpath,{libdir}"], check=True)
    L = refmex.lib(so)
    c = synth.make_correspondences(4900, N=150, outlier_ratio=0.3)
    match = np.stack([np.arange(150) + 1, np.random.default_rng(4).permutation(150) + 1], 1).astype(np.float64)
    draws = synth.make_draws(4901, 700, 150)

    def call(args, nout):
        keep = [np.ascontiguousarray(a, np.float64) for a in args]
        ins = [L.stub_wrap(6, a.shape[1] if a.ndim == 2 else 1, a.shape[0], a.ctypes.data) for a in keep]
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
            res.append(np.ctypeslib.as_array(C.cast(m.data, C.POINTER(C.c_double)), shape=(m.n, m.m)).copy().T
                       if m.cls == 6 and m.m * m.n else np.zeros((m.m, m.n)))
            L.mxDestroyArray(out_arr[i])
        return res

    # arrays are passed as (columns, rows) C-contiguous = MATLAB column-major rows x columns
    rot, trans, sta, op_num, good = call([c.Ya, c.Yb, match, (draws + 1).astype(np.float64)], 5)
    o = orc.vodometry_dr_ye(c.Ya, c.Yb, samples=draws)
    assert sta[0, 0] == o.state == 1 and op_num[0, 0] == o.op_num
    np.testing.assert_array_equal(good.ravel().astype(int) - 1, np.flatnonzero(o.mask))
    assert rn.rot_angle(rot, o.R) < 1e-9 and np.abs(trans.ravel() - o.T).max() < 1e-9
    # seeded form: a scalar in place of the draws
    rot2, _, sta2, op2 = call([c.Ya, c.Yb, match, np.array([[5.0]])], 4)
    o2 = orc.vodometry_dr_ye(c.Ya, c.Yb, match=match.astype(np.int32), seed=5, pair=0)
    assert op2[0, 0] == o2.op_num and rn.rot_angle(rot2, o2.R) < 1e-9
    with pytest.raises(refmex.MexError, match="same size"):
        call([c.Ya, c.Yb[:10], match], 1)
    with pytest.raises(refmex.MexError, match="required"):
        call([c.Ya, c.Yb], 1)
    # fewer than 4 matches: SolutionState 4, no error (vodometry_dr_ye.m:152-160)
    _, _, sta3, op3 = call([c.Ya[:3], c.Yb[:3], match[:3]], 4)
    assert sta3[0, 0] == 4 and op3[0, 0] == 0

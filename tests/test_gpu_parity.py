"""GPU parity suite (-m gpu): libpre3.so through its C ABI vs the CPU oracle on the same
seeded inputs.  Bar: bit-exact NN indices, scores, fit states, cardinalities, inlier masks and
selected hypothesis; refit rotation within 1e-9 rad and translation within 1e-9 m (the refit
sums are tree-reduced on the GPU, sequential in the oracle).
"""
import importlib

import numpy as np
import pytest

from oracle import ref_numpy as rn

pytestmark = pytest.mark.gpu

TOL_ROT = 1e-9  # rad
TOL_T = 1e-9    # m


# ------------------------------------------------------------------------------------------
# stage 1
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["f64", "f32", "u8", "i8", "k2one", "ties", "knn"])
@pytest.mark.parametrize("engine", [0, 1])
def test_siftmatch_vs_reference_golden(ctx, golden, name, engine):
    ctx.set_match_engine(engine)
    try:
        pairs, score = ctx.siftmatch(golden[f"{name}_L1"], golden[f"{name}_L2"], float(golden[f"{name}_thresh"]))
    finally:
        ctx.set_match_engine(0)
    np.testing.assert_array_equal(pairs.T + 1, golden[f"{name}_matches"].astype(np.int64))
    np.testing.assert_array_equal(score, golden[f"{name}_D"])


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.uint8, np.int8])
@pytest.mark.parametrize("engine", [0, 1])
def test_siftmatch_vs_oracle(ctx, orc, synth, dtype, engine):
    ctx.set_match_engine(engine)
    try:
        for seed, (K1, K2) in enumerate([(512, 512), (300, 517), (1, 40), (129, 1), (700, 2048)]):
            fp = synth.make_frame_pair(40 + seed, K1=K1, K2=K2, n_corr=min(K1, K2) // 2)
            d1, d2 = fp.desc1, fp.desc2
            if dtype == np.float32:
                d1, d2 = d1.astype(np.float32), d2.astype(np.float32)
            elif dtype == np.uint8:
                d1, d2 = synth.to_uint8(d1), synth.to_uint8(d2)
            elif dtype == np.int8:
                d1, d2 = (synth.to_uint8(d1) // 2).astype(np.int8), (synth.to_uint8(d2) // 2).astype(np.int8)
            if K2 > 10:
                d2[K2 - 1] = d2[3]  # exact duplicate column: first index wins, ratio test fails unless best == 0
            pairs, score = ctx.siftmatch(d1, d2, 1.5)
            op, os_ = orc.siftmatch(d1, d2, 1.5)
            np.testing.assert_array_equal(pairs, op)
            np.testing.assert_array_equal(score, os_)
            # nout == 1 form: rows certified by the proposal brackets skip the exact distance
            pairs1, none = ctx.siftmatch(d1, d2, 1.5, want_score=False)
            assert none is None
            np.testing.assert_array_equal(pairs1, op)
            if np.issubdtype(dtype, np.floating) and K1 >= 300:
                assert len(pairs) >= 0.4 * min(K1, K2)  # the planted matches are found
    finally:
        ctx.set_match_engine(0)


def test_siftmatch_adversarial_ratio_band(ctx, orc):
    """Rows whose second/best ratio sits exactly at / next to the threshold, identical rows
    (distance 0), and near-duplicates: the tensor-core proposal must hand these to the exact path."""
    rng = np.random.default_rng(5)
    base = np.abs(rng.normal(size=(64, 128)))
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    L1 = base.astype(np.float32).astype(np.float64)
    L2 = np.concatenate([L1, L1 + 1e-4 * rng.normal(size=L1.shape), L1[::-1] * (1 + 1e-7)])
    L2 = L2.astype(np.float32).astype(np.float64)
    for thresh in (1.5, 1.0, 1.0000001, 3.0):
        pairs, score = ctx.siftmatch(L1, L2, thresh)
        op, os_ = orc.siftmatch(L1, L2, thresh)
        np.testing.assert_array_equal(pairs, op)
        np.testing.assert_array_equal(score, os_)
        np.testing.assert_array_equal(ctx.siftmatch(L1, L2, thresh, want_score=False)[0], op)


def test_siftmatch_batch_ragged(ctx, orc, synth):
    P, K1, K2 = 5, 256, 384
    L1 = np.zeros((P, K1, 128))
    L2 = np.zeros((P, K2, 128))
    k1c = np.array([256, 100, 0, 1, 255], np.int32)
    k2c = np.array([384, 1, 50, 0, 383], np.int32)
    for p in range(P):
        fp = synth.make_frame_pair(70 + p, K1=K1, K2=K2, n_corr=90)
        L1[p], L2[p] = fp.desc1, fp.desc2
    out = ctx.siftmatch_batch(L1, L2, 1.5, k1c, k2c)
    for p in range(P):
        op, os_ = orc.siftmatch(L1[p, : k1c[p]], L2[p, : k2c[p]], 1.5)
        np.testing.assert_array_equal(out[p][0], op)
        np.testing.assert_array_equal(out[p][1], os_)


# ------------------------------------------------------------------------------------------
# stage 2
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", [0, 1])
@pytest.mark.parametrize("k", [3, 4, 5, 8])
def test_fit_batch_bit_exact(ctx, orc, synth, method, k):
    c = synth.make_correspondences(11 + k, N=300, outlier_ratio=0.3)
    samples = synth.make_samples(12 + k, 1500, 300, k)
    # a few degenerate sets: repeated index (rank-deficient), collinear points
    samples[0] = np.array([7] * k)
    R, T, st = ctx.fit_batch(c.Ya, c.Yb, samples, method)
    for h in range(len(samples)):
        if method == 0:
            r0, t0, s0 = orc.find_transform_matrix(c.Ya, c.Yb, samples[h])
        else:
            _, r0, t0, _ = orc.horn(c.Yb, c.Ya, 0, samples[h], allow_small=True)
            s0 = 1
        assert st[h] == s0
        np.testing.assert_array_equal(R[h], r0)
        np.testing.assert_array_equal(T[h], t0)


def test_full_set_fits(ctx, orc):
    rng = np.random.default_rng(2)
    for n in (4, 5, 37, 1000):
        Yb = rng.normal(size=(n, 3))
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        Ya = 1.3 * Yb @ R.T + rng.normal(size=3) + 1e-3 * rng.normal(size=(n, 3))
        rot, tr, st = ctx.find_transform_matrix(Ya, Yb)
        r0, t0, s0 = orc.find_transform_matrix(Ya, Yb)
        assert st == s0
        np.testing.assert_array_equal(rot, r0)
        np.testing.assert_array_equal(tr, t0)
        for do_scale in (0, 1):
            s, Rh, Th, err = ctx.horn(Yb, Ya, do_scale)
            s0, R0, T0, e0 = orc.horn(Yb, Ya, do_scale)
            assert s == s0 and err == e0
            np.testing.assert_array_equal(Rh, R0)
            np.testing.assert_array_equal(Th, T0)
            s1, R1, T1, e1 = rn.horn(Yb, Ya, bool(do_scale))
            assert rn.rot_angle(Rh, R1) < TOL_ROT and np.abs(Th - T1).max() < TOL_T


# ------------------------------------------------------------------------------------------
# stage 3
# ------------------------------------------------------------------------------------------
def test_score_batch_bit_exact(ctx, orc, synth):
    c = synth.make_correspondences(21, N=1111, outlier_ratio=0.4)
    samples = synth.make_samples(22, 600, 1111, 5)
    R, T, st = ctx.fit_batch(c.Ya, c.Yb, samples, 0)
    thr = orc.distance_threshold(c.Yb)
    cnt, es, mk = ctx.score_batch(R, T, c.Ya, c.Yb, thr)
    for h in range(len(samples)):
        c0, m0, e0 = orc.score(R[h], T[h], c.Ya, c.Yb, thr)
        assert cnt[h] == c0 and es[h] == e0
        np.testing.assert_array_equal(mk[h], m0)


def test_score_borderline_residuals(ctx, orc):
    """Residuals placed within a few ulps of the threshold: fp32 scoring alone would misclassify
    them; the fp64 recheck must reproduce the reference's strict '<'."""
    rng = np.random.default_rng(9)
    N = 4096
    Yb = rng.uniform(-3, 3, size=(N, 3))
    R = np.eye(3)
    T = np.array([0.01, -0.02, 0.03])
    thr = 0.0125
    u = rng.normal(size=(N, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    scale = thr * (1 + rng.choice([-1e-15, -1e-12, -1e-9, -1e-7, 0.0, 1e-16, 1e-12, 1e-9, 1e-7], size=N))
    Ya = Yb + T - u * scale[:, None]
    cnt, es, mk = ctx.score_batch(R[None], T[None], Ya, Yb, thr)
    c0, m0, e0 = orc.score(R, T, Ya, Yb, thr)
    assert 0.2 * N < c0 < 0.8 * N
    assert cnt[0] == c0 and es[0] == e0
    np.testing.assert_array_equal(mk[0], m0)


# ------------------------------------------------------------------------------------------
# stages 2-4: the RANSAC loop (config 1: N = 300, 30 % outliers, 2000 sample sets)
# ------------------------------------------------------------------------------------------
def _check_ransac(g, o, N):
    assert g.status == o.status
    if o.status != 0:
        return
    assert g.thr == o.thr
    assert (g.best_fit, g.best_sample, g.best_iter, g.n_iter, g.n_consumed) == \
        (o.best_fit, o.best_sample, o.best_iter, o.n_iter, o.n_consumed)
    np.testing.assert_array_equal(g.mask, o.mask)
    assert g.error_sum == o.error_sum
    np.testing.assert_array_equal(g.R_hyp, o.R_hyp)
    np.testing.assert_array_equal(g.T_hyp, o.T_hyp)
    assert g.state == o.state
    if o.state in (1, 2):
        assert rn.rot_angle(g.R, o.R) < TOL_ROT and np.abs(g.T - o.T).max() < TOL_T
    else:  # rot = H (not a rotation), trans = 0 (find_transform_matrix.m:35-36,:40-41)
        np.testing.assert_allclose(g.R, o.R, rtol=0, atol=1e-12)
        np.testing.assert_array_equal(g.T, o.T)
    if g.counts is not None and o.counts is not None:
        np.testing.assert_array_equal(g.counts, o.counts)
        np.testing.assert_array_equal(g.states[: o.n_consumed], o.states[: o.n_consumed])


@pytest.mark.parametrize("method,k", [(0, 5), (0, 3), (1, 5), (1, 4), (1, 3)])
@pytest.mark.parametrize("adaptive", [True, False])
def test_ransac_config1(ctx, orc, synth, pre3, method, k, adaptive):
    for seed in range(6):
        c = synth.make_correspondences(1000 + seed, N=300, outlier_ratio=0.30)
        samples = synth.make_samples(2000 + seed, 2000, 300, k)
        opts = pre3.make_opts(method=method, k=k, max_iteration=2000, adaptive=adaptive, distance_threshold=0.012)
        g = ctx.ransac(c.Ya, c.Yb, samples, opts)
        o = orc.ransac(c.Ya, c.Yb, samples, method=method, max_iteration=2000, distance_threshold=0.012,
                       adaptive=adaptive)
        _check_ransac(g, o, 300)
        assert rn.rot_angle(g.R, c.R) < 2e-3 and np.abs(g.T - c.t).max() < 5e-3


def test_ransac_seeded_samples_match_oracle_generator(ctx, orc, synth, pre3):
    c = synth.make_correspondences(77, N=257, outlier_ratio=0.5)
    for k in (3, 5):
        opts = pre3.make_opts(method=0, k=k, max_iteration=500, adaptive=True, H=500, seed=424242)
        g = ctx.ransac(c.Ya, c.Yb, None, opts)
        samples = orc.sample_sets(424242, 0, 500, 257, k)
        o = orc.ransac(c.Ya, c.Yb, samples, method=0, max_iteration=500, adaptive=True)
        _check_ransac(g, o, 257)


def test_ransac_edge_cases(ctx, orc, synth, pre3):
    c = synth.make_correspondences(5, N=120, outlier_ratio=0.2)
    samples = synth.make_samples(6, 50, 120, 5)
    # fewer correspondences than k
    g = ctx.ransac(c.Ya[:4], c.Yb[:4], samples % 4, pre3.make_opts(adaptive=False))
    assert g.status == 1
    # MaxIteration bound, perfect data (card == N stops the loop), degenerate first sample, all outliers
    for Ya, Yb, s, kw in [
        (c.Ya, c.Yb, samples, dict(max_iteration=10, adaptive=False)),
        (c.Yb @ c.R.T + c.t, c.Yb, samples, dict(adaptive=True)),
        (np.random.default_rng(1).normal(size=(120, 3)), c.Yb, samples, dict(adaptive=True)),
        (c.Ya, c.Yb, samples[:1], dict(max_iteration=1, adaptive=True)),  # no iteration runs -> status 2
    ]:
        g = ctx.ransac(Ya, Yb, s, pre3.make_opts(**kw))
        o = orc.ransac(Ya, Yb, s, max_iteration=kw.get("max_iteration", 2000), adaptive=kw["adaptive"])
        _check_ransac(g, o, len(Ya))
    Yb = c.Yb.copy()
    Yb[:5] = np.outer(np.arange(5.0), [1.0, 1.0, 1.0])
    Ya = Yb @ c.R.T + c.t
    s2 = samples.copy()
    s2[0] = np.arange(5)
    g = ctx.ransac(Ya, Yb, s2, pre3.make_opts(adaptive=False, max_iteration=20))
    o = orc.ransac(Ya, Yb, s2, adaptive=False, max_iteration=20)
    assert o.states[0] == -1
    _check_ransac(g, o, 120)


def test_ransac_batch_ragged(ctx, orc, synth, pre3):
    P, Nmax, H, k = 9, 320, 600, 5
    Ya = np.zeros((P, Nmax, 3))
    Yb = np.zeros((P, Nmax, 3))
    n = np.array([320, 300, 5, 4, 0, 17, 319, 128, 64], np.int32)
    samples = np.zeros((P, H, k), np.int32)
    for p in range(P):
        c = synth.make_correspondences(300 + p, N=max(int(n[p]), 1), outlier_ratio=0.3)
        Ya[p, : n[p]], Yb[p, : n[p]] = c.Ya[: n[p]], c.Yb[: n[p]]
        if n[p] >= k:
            samples[p] = synth.make_samples(400 + p, H, int(n[p]), k)
    opts = pre3.make_opts(method=0, k=k, max_iteration=2000, adaptive=True)
    res, masks = ctx.ransac_batch(Ya, Yb, n, samples, opts)
    for p in range(P):
        g = pre3.unpack_result(res[p], masks[p, : n[p]].astype(bool))
        o = orc.ransac(Ya[p, : n[p]], Yb[p, : n[p]], samples[p], method=0, max_iteration=2000, adaptive=True)
        _check_ransac(g, o, n[p])
        assert not masks[p, n[p]:].any()


# ------------------------------------------------------------------------------------------
# whole pairs: match -> gather -> RANSAC (configs 1 / 3 shape)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cls", ["f64", "f32"])
def test_pairs_vs_oracle(ctx, orc, synth, pre3, cls):
    P, K = 6, 512
    fps = [synth.make_frame_pair(900 + p, K1=K, K2=K, n_corr=300, outlier_ratio=0.30) for p in range(P)]
    d1 = np.stack([f.desc1 for f in fps])
    d2 = np.stack([f.desc2 for f in fps])
    x1 = np.stack([f.xyz1 for f in fps])
    x2 = np.stack([f.xyz2 for f in fps])
    if cls == "f32":
        d1, d2 = d1.astype(np.float32), d2.astype(np.float32)
    opts = pre3.make_opts(method=0, k=5, max_iteration=2000, adaptive=True, H=2000, seed=99)
    res, matches, masks = ctx.pairs(d1, d2, x1, x2, opts, pair_id0=1000)
    for p in range(P):
        # float32-valued doubles: the f32 and f64 classes give the same matches here only if the
        # accumulation type does not change a decision; the oracle is run in the same class
        om, o = orc.pair(d1[p].astype(np.float64), d2[p].astype(np.float64), x1[p], x2[p], 99, 1000 + p, H=2000) \
            if cls == "f64" else (None, None)
        n = int(res["n_matches"][p])
        if cls == "f32":
            op, _ = orc.siftmatch(d1[p], d2[p], 1.5)
            np.testing.assert_array_equal(matches[p, :n], op)
            Ya, Yb = x1[p][op[:, 0]], x2[p][op[:, 1]]
            o = orc.ransac(Ya, Yb, orc.sample_sets(99, 1000 + p, 2000, len(op), 5), method=0, max_iteration=2000)
        else:
            np.testing.assert_array_equal(matches[p, :n], om)
        g = pre3.unpack_result(res[p], masks[p, :n].astype(bool))
        _check_ransac(g, o, n)
        assert n >= 280 and rn.rot_angle(g.R, fps[p].R) < 2e-3 and np.abs(g.T - fps[p].t).max() < 5e-3


def test_pairs_chunked_host_path_equals_device_path(ctx, synth, pre3):
    """pre3_pairs (host buffers, chunked + double-buffered) == pre3_pairs_dev (resident buffers)."""
    import torch
    b = synth.make_batch_torch(40, 5, "cuda", K1=256, K2=256, n_corr=150)
    opts = pre3.make_opts(method=0, k=5, max_iteration=500, adaptive=True, H=500, seed=3)
    res_d = torch.zeros(40, 240, dtype=torch.uint8, device="cuda")
    m_d = torch.zeros(40, 256, 2, dtype=torch.int32, device="cuda")
    k_d = torch.zeros(40, 256, dtype=torch.uint8, device="cuda")
    ctx.pairs_dev(b["desc1"], b["desc2"], b["xyz1"], b["xyz2"], opts, res_d, m_d, k_d, pair_id0=0)
    ctx.sync()
    res, matches, masks = ctx.pairs(b["desc1"].cpu().numpy(), b["desc2"].cpu().numpy(), b["xyz1"].cpu().numpy(),
                                    b["xyz2"].cpu().numpy(), opts, pair_id0=0)
    rd = np.frombuffer(res_d.cpu().numpy().tobytes(), dtype=pre3.RESULT_DTYPE)
    assert rd.tobytes() == res.tobytes()
    for p in range(40):
        n = int(res["n_matches"][p])
        np.testing.assert_array_equal(matches[p, :n], m_d[p, :n].cpu().numpy())
        np.testing.assert_array_equal(masks[p, :n], k_d[p, :n].cpu().numpy())


# ------------------------------------------------------------------------------------------
# full-size properties (configs 2, 3, 5 shapes) -- no oracle at these sizes
# ------------------------------------------------------------------------------------------
def test_pairs_host_narrowing_is_lossless_and_falls_back(synth, pre3, monkeypatch):
    """Host path: float-exact double descriptors cross PCIe as float (half the bytes), others as double;
    either way the results equal the device path bit for bit."""
    import torch
    monkeypatch.setenv("PRE3_HOST_F32", "1")  # the default depends on the host's core count per rank
    ctx = pre3.Context(0)
    P, K = 6, 256
    b = synth.make_batch_torch(P, 910, "cuda", K1=K, K2=K, n_corr=150)
    opts = pre3.make_opts(H=500, seed=3)
    for exact in (True, False):
        d1, d2 = b["desc1"].clone(), b["desc2"].clone()
        if not exact:  # full double precision in one value of pair 3
            d1[3, 7, 5] += 1e-13
        res = torch.zeros(P, 240, dtype=torch.uint8, device="cuda")
        m = torch.zeros(P, K, 2, dtype=torch.int32, device="cuda")
        ctx.pairs_dev(d1, d2, b["xyz1"], b["xyz2"], opts, res, m)
        ctx.sync()
        dev = np.frombuffer(res.cpu().numpy().tobytes(), dtype=pre3.RESULT_DTYPE)
        t0 = ctx.transfer_bytes()[0]
        hres, hm, _ = ctx.pairs(d1.cpu().numpy(), d2.cpu().numpy(), b["xyz1"].cpu().numpy(), b["xyz2"].cpu().numpy(), opts)
        moved = ctx.transfer_bytes()[0] - t0
        assert hres.tobytes() == dev.tobytes()
        md = m.cpu().numpy()
        for p in range(P):  # entries beyond the pair's match count are not written
            n = int(dev["n_matches"][p])
            np.testing.assert_array_equal(hm[p, :n], md[p, :n])
        desc_bytes = 2 * P * K * 128 * 8
        assert moved < 0.6 * desc_bytes if exact else moved > desc_bytes
    ctx.close()


@pytest.mark.parametrize("cls", ["f64", "f32", "u8"])
@pytest.mark.parametrize("engine", [0, 1])
def test_sequence_equals_pairs(ctx, orc, synth, pre3, cls, engine):
    """F consecutive frames through pre3_sequence(_dev) == the F-1 pairs (frame p, frame p+1) through pre3_pairs_dev
    (bit for bit: records, matches, masks), ragged frame sizes included; pair 0 also against the oracle."""
    import torch
    F, K = 9, 200  # K not a multiple of 128; 8 pairs
    sq = synth.make_sequence_torch(F, 77, "cuda", K=K, n_corr=120)
    desc, xyz = sq["desc"], sq["xyz"]
    if cls == "f32":
        desc = desc.to(torch.float32)
    elif cls == "u8":
        desc = torch.clamp(torch.floor(512.0 * desc + 0.5), 0, 255).to(torch.uint8)
    kc = torch.tensor([K, K - 3, K, K - 17, K, K, K - 1, K, K], dtype=torch.int32, device="cuda")
    opts = pre3.make_opts(H=400, seed=9)
    ctx.set_match_engine(engine)
    try:
        P = F - 1
        r1 = torch.zeros(P, 240, dtype=torch.uint8, device="cuda")
        m1 = torch.zeros(P, K, 2, dtype=torch.int32, device="cuda")
        k1 = torch.zeros(P, K, dtype=torch.uint8, device="cuda")
        ctx.pairs_dev(desc[:-1].contiguous(), desc[1:].contiguous(), xyz[:-1].contiguous(), xyz[1:].contiguous(), opts,
                      r1, m1, k1, pair_id0=5, k1_count=kc[:-1].contiguous(), k2_count=kc[1:].contiguous())
        r2 = torch.zeros(P, 240, dtype=torch.uint8, device="cuda")
        m2 = torch.zeros(P, K, 2, dtype=torch.int32, device="cuda")
        k2 = torch.zeros(P, K, dtype=torch.uint8, device="cuda")
        ctx.sequence_dev(desc, xyz, opts, r2, m2, k2, pair_id0=5, k_count=kc)
        ctx.sync()
        a = np.frombuffer(r1.cpu().numpy().tobytes(), dtype=pre3.RESULT_DTYPE)
        b = np.frombuffer(r2.cpu().numpy().tobytes(), dtype=pre3.RESULT_DTYPE)
        assert a.tobytes() == b.tobytes() and (a["status"] == 0).all() and (a["best_fit"] > 40).all()
        hres, hm, hk = ctx.sequence(desc.cpu().numpy(), xyz.cpu().numpy(), opts, pair_id0=5, k_count=kc.cpu().numpy())
        assert hres.tobytes() == a.tobytes()
        for p in range(P):
            n = int(a["n_matches"][p])
            np.testing.assert_array_equal(m1[p, :n].cpu().numpy(), m2[p, :n].cpu().numpy())
            np.testing.assert_array_equal(hm[p, :n], m2[p, :n].cpu().numpy())
            np.testing.assert_array_equal(k1[p, :n].cpu().numpy(), k2[p, :n].cpu().numpy())
            np.testing.assert_array_equal(hk[p, :n], k2[p, :n].cpu().numpy())
        if cls == "f64":
            d0, d1 = desc[0].cpu().numpy(), desc[1, : K - 3].cpu().numpy()
            om, o = orc.pair(d0, d1, xyz[0].cpu().numpy(), xyz[1, : K - 3].cpu().numpy(), 9, 5, H=400)
            np.testing.assert_array_equal(hm[0, : len(om)], om)
            assert (o.best_fit, o.best_sample) == (a["best_fit"][0], a["best_sample"][0])
        # a single frame has no pair
        res0, _, _ = ctx.sequence(desc[:1].cpu().numpy(), xyz[:1].cpu().numpy(), opts)
        assert len(res0) == 0
    finally:
        ctx.set_match_engine(0)


@pytest.mark.parametrize("cls", ["f64", "u8"])
@pytest.mark.parametrize("graphs", [False, True])
def test_pipeline_equals_unchunked(pre3, synth, cls, graphs):
    """pre3_set_pipeline: the chunked, four-stream form of pre3_sequence_dev / pre3_pairs_dev gives the bytes of the
    single-stream call (records, matches, masks) -- eager and as a replayed CUDA graph, ragged frame sizes included,
    chunk boundaries that do not divide the pair count; one pair is also compared with the oracle elsewhere
    (test_sequence_equals_pairs), so this pins the pipeline to the already-pinned path."""
    import torch
    ctx = pre3.Context(0)
    try:
        F, K = 418, 160  # 417 pairs -> 3 chunks of 139 pairs
        sq = synth.make_sequence_torch(F, 901, "cuda", K=K, n_corr=100)
        desc, xyz = sq["desc"], sq["xyz"]
        if cls == "u8":
            desc = torch.clamp(torch.floor(512.0 * desc + 0.5), 0, 255).to(torch.uint8)
        kc = torch.full((F,), K, dtype=torch.int32, device="cuda")
        kc[7] = K - 5
        kc[139] = K - 31  # the frame on a chunk boundary
        kc[300] = K - 1
        opts = pre3.make_opts(H=300, seed=4)
        P = F - 1

        def run(chunks, seq):
            ctx.set_pipeline(chunks)
            r = torch.zeros(P, 240, dtype=torch.uint8, device="cuda")
            m = torch.zeros(P, K, 2, dtype=torch.int32, device="cuda")
            k = torch.zeros(P, K, dtype=torch.uint8, device="cuda")
            for _ in range(3 if graphs else 1):  # eager, capture, replay
                r.zero_(); m.zero_(); k.zero_()
                torch.cuda.synchronize()
                if seq:
                    ctx.sequence_dev(desc, xyz, opts, r, m, k, pair_id0=11, k_count=kc)
                else:
                    ctx.pairs_dev(d1, d2, x1, x2, opts, r, m, k, pair_id0=11, k1_count=kc1, k2_count=kc2)
                ctx.sync()
            return r.cpu().numpy().tobytes(), m.cpu().numpy(), k.cpu().numpy()

        d1, d2 = desc[:-1].contiguous(), desc[1:].contiguous()
        x1, x2 = xyz[:-1].contiguous(), xyz[1:].contiguous()
        kc1, kc2 = kc[:-1].contiguous(), kc[1:].contiguous()
        ctx.set_graphs(graphs)
        for seq in (True, False):
            r0, m0, k0 = run(0, seq)
            rec = np.frombuffer(r0, dtype=pre3.RESULT_DTYPE)
            assert (rec["status"] == 0).all() and (rec["best_fit"] > 30).all()
            for chunks in (3, 2):
                r1, m1, k1 = run(chunks, seq)
                assert r1 == r0
                for p in range(P):
                    n = int(rec["n_matches"][p])
                    np.testing.assert_array_equal(m1[p, :n], m0[p, :n])
                    np.testing.assert_array_equal(k1[p, :n], k0[p, :n])
        with pytest.raises(Exception):
            ctx.set_pipeline(65)
    finally:
        ctx.close()


@pytest.mark.parametrize("cls", ["f64", "f32"])
@pytest.mark.parametrize("K,P", [(512, 160), (500, 37), (300, 3)])
def test_fused_sequence_matcher_equals_separate_kernels(pre3, synth, cls, K, P):
    """k_tc_seq_fused (conversion + proposal GEMM of a sequence in one kernel, frames resident in shared memory) against
    k_tc_convert + k_tc_gemm_pair (PRE3_TC_FUSED=0): records, matches and masks bit for bit; more pairs than CTA pairs
    (160 > 74: runs of 2-3 pairs per CTA pair, every buffer rotation), fewer (37, 3), ragged frames, K < 512."""
    import os
    import torch
    F = P + 1
    sq = synth.make_sequence_torch(F, 313 + K, "cuda", K=K, n_corr=min(250, K // 2))
    desc, xyz = sq["desc"], sq["xyz"]
    if cls == "f32":
        desc = desc.to(torch.float32)
    kc = torch.full((F,), K, dtype=torch.int32, device="cuda")
    kc[1] = K - 7
    kc[F - 1] = K - 40
    if F > 20:
        kc[17] = K - 129
    opts = pre3.make_opts(H=300, seed=21)
    out = {}
    old = os.environ.get("PRE3_TC_FUSED")
    try:
        for mode in ("1", "0"):
            os.environ["PRE3_TC_FUSED"] = mode
            ctx = pre3.Context(0)
            ctx.set_match_engine(1)
            r = torch.zeros(P, 240, dtype=torch.uint8, device="cuda")
            m = torch.zeros(P, K, 2, dtype=torch.int32, device="cuda")
            k = torch.zeros(P, K, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            ctx.sequence_dev(desc, xyz, opts, r, m, k, pair_id0=3, k_count=kc)
            ctx.sync()
            out[mode] = (r.cpu().numpy().tobytes(), m.cpu().numpy(), k.cpu().numpy())
            ctx.close()
    finally:
        if old is None:
            os.environ.pop("PRE3_TC_FUSED", None)
        else:
            os.environ["PRE3_TC_FUSED"] = old
    rec = np.frombuffer(out["1"][0], dtype=pre3.RESULT_DTYPE)
    assert out["1"][0] == out["0"][0]
    assert (rec["status"] == 0).all() and (rec["n_matches"] > 50).all()
    for p in range(P):
        n = int(rec["n_matches"][p])
        np.testing.assert_array_equal(out["1"][1][p, :n], out["0"][1][p, :n])
        np.testing.assert_array_equal(out["1"][2][p, :n], out["0"][2][p, :n])


@pytest.mark.parametrize("scale", [1.0, 512.0])
@pytest.mark.parametrize("ratio", [1.5, 1.0000001])
def test_fused_sequence_matcher_adversarial_vs_reference(pre3, orc, scale, ratio):
    """The matches of k_tc_seq_fused (fp16 proposal, |x|^2 summed in fp32, certified rescore) against the reference
    siftmatch on frames built to sit on the decisions: a frame holds exact copies of the previous frame's descriptors,
    copies perturbed by 1e-4 and 1e-7 (best / second best a few ulps of the proposal apart, ratios at the threshold),
    duplicated columns (first index wins), zero rows; at unit norm and at the uint8 scale (integer-valued doubles up to
    ~150, |x|^2 ~ 2.6e5, where the fp32 norm is off by up to 0.25 absolute)."""
    import torch
    rng = np.random.default_rng(77)
    F, K = 6, 512
    base = np.abs(rng.normal(size=(K, 128)))
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    frames = [base]
    for f in range(1, F):
        prev = frames[-1]
        cur = np.abs(rng.normal(size=(K, 128)))
        cur /= np.linalg.norm(cur, axis=1, keepdims=True)
        perm = rng.permutation(K)
        cur[perm[:128]] = prev[:128]                                             # exact copies (distance 0)
        cur[perm[128:256]] = prev[128:256] + 1e-4 * rng.normal(size=(128, 128))  # near copies
        cur[perm[256:320]] = prev[:64] * (1 + 1e-7)                              # second copy of rows that already have one
        cur[perm[320:336]] = cur[perm[0:16]]                                     # duplicated columns
        cur[perm[336:340]] = 0.0
        frames.append(cur)
    desc = np.stack(frames)
    if scale != 1.0:
        desc = np.rint(desc * scale)
    desc = desc.astype(np.float32).astype(np.float64)
    xyz = rng.normal(size=(F, K, 3)) + np.array([0.0, 0.0, 3.0])
    P = F - 1
    opts = pre3.make_opts(H=64, ratio=ratio, seed=5)
    ctx = pre3.Context(0)
    try:
        ctx.set_match_engine(1)
        r = torch.zeros(P, 240, dtype=torch.uint8, device="cuda")
        m = torch.zeros(P, K, 2, dtype=torch.int32, device="cuda")
        k = torch.zeros(P, K, dtype=torch.uint8, device="cuda")
        ctx.sequence_dev(torch.from_numpy(desc).cuda(), torch.from_numpy(xyz).cuda(), opts, r, m, k)
        ctx.sync()
        rec = np.frombuffer(r.cpu().numpy().tobytes(), dtype=pre3.RESULT_DTYPE)
        mm = m.cpu().numpy()
    finally:
        ctx.close()
    for p in range(P):
        op, _ = orc.siftmatch(desc[p], desc[p + 1], ratio)
        n = int(rec["n_matches"][p])
        assert n == len(op) and n >= 64  # at least the rows with a single exact copy are matched
        np.testing.assert_array_equal(mm[p, :n], op)


def test_graph_replay_equals_eager(pre3, synth):
    """pre3_set_graphs: the captured launch sequence of a repeated pre3_sequence_dev signature gives the same bytes as
    the eager calls, also after the inputs behind the same pointers changed; a new signature falls back to eager."""
    import torch
    ctx = pre3.Context(0)   # its own (non-default) stream: the legacy default stream cannot be captured
    F, K = 33, 256
    sq = synth.make_sequence_torch(F, 512, "cuda", K=K, n_corr=150)
    sq2 = synth.make_sequence_torch(F, 513, "cuda", K=K, n_corr=150)
    opts = pre3.make_opts(H=500, seed=3)
    P = F - 1
    torch.cuda.synchronize()

    def run(desc, xyz):
        r = torch.zeros(P, 240, dtype=torch.uint8, device="cuda")
        m = torch.zeros(P, K, 2, dtype=torch.int32, device="cuda")
        k = torch.zeros(P, K, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ctx.sequence_dev(desc, xyz, opts, r, m, k, pair_id0=7)
        ctx.sync()
        return r.cpu().numpy().tobytes(), k.cpu().numpy().tobytes()

    eager1, eager2 = run(sq["desc"], sq["xyz"]), run(sq2["desc"], sq2["xyz"])
    ctx.set_graphs(True)
    desc, xyz = sq["desc"].clone(), sq["xyz"].clone()
    r = torch.zeros(P, 240, dtype=torch.uint8, device="cuda")
    m = torch.zeros(P, K, 2, dtype=torch.int32, device="cuda")
    k = torch.zeros(P, K, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    outs = []
    launches = []
    for it in range(5):
        if it == 3:  # new inputs behind the same pointers: the replayed graph must pick them up
            desc.copy_(sq2["desc"]); xyz.copy_(sq2["xyz"])
            torch.cuda.synchronize()
        l0 = ctx.launch_count()
        ctx.sequence_dev(desc, xyz, opts, r, m, k, pair_id0=7)
        ctx.sync()
        launches.append(ctx.launch_count() - l0)
        outs.append((r.cpu().numpy().tobytes(), k.cpu().numpy().tobytes()))
    assert outs[0] == outs[1] == outs[2] == eager1 and outs[3] == outs[4] == eager2
    assert len(set(launches)) == 1 and launches[0] > 5     # a replay stands for the same number of kernel launches
    # a different signature (other pair_id0) runs eagerly and leaves the cached graph alone
    ctx.sequence_dev(desc, xyz, opts, r, m, k, pair_id0=8)
    ctx.sync()
    ctx.set_graphs(False)
    ctx.close()


def test_matching_full_size_properties(ctx, synth):
    """2k x 2k descriptors: planted matches are found; matching L against itself returns the
    identity with score 0; results are independent of the batch position."""
    import torch
    b = synth.make_batch_torch(4, 11, "cuda", K1=2048, K2=2048, n_corr=1024)
    pairs = torch.zeros(4, 2048, 2, dtype=torch.int32, device="cuda")
    score = torch.zeros(4, 2048, dtype=torch.float64, device="cuda")
    n = torch.zeros(4, dtype=torch.int32, device="cuda")
    ctx.siftmatch_batch_dev(b["desc1"], b["desc2"], pairs, score, n)
    ctx.sync()
    assert (n.cpu().numpy() >= 1000).all()
    ctx.siftmatch_batch_dev(b["desc1"], b["desc1"], pairs, score, n)
    ctx.sync()
    assert (n.cpu().numpy() == 2048).all()
    ar = torch.arange(2048, dtype=torch.int32, device="cuda")
    assert (pairs[..., 0] == ar).all() and (pairs[..., 1] == ar).all() and (score == 0).all()


def test_stress_pair_properties(ctx, synth, pre3):
    """20k correspondences, 60 % outliers, 100k seeded hypotheses (config 5 at 1/10 of the
    hypotheses): the winner's mask is the planted inlier set up to the noise tail, and the
    refit recovers the planted motion."""
    c = synth.make_correspondences(5005, N=20000, outlier_ratio=0.60)
    opts = pre3.make_opts(method=0, k=5, max_iteration=100001, adaptive=False, H=100000, seed=1)
    g = ctx.ransac(c.Ya, c.Yb, None, opts)
    assert g.status == 0 and g.n_consumed == 100000  # (sets giving a reflection are skipped uncounted)
    assert g.n_iter == int((g.states != -1).sum())
    assert g.best_fit == g.mask.sum() == g.counts.max()
    assert (g.mask & ~c.inlier).sum() <= 20 and (g.mask & c.inlier).sum() >= 0.8 * c.inlier.sum()
    assert rn.rot_angle(g.R, c.R) < 1e-3 and np.abs(g.T - c.t).max() < 2e-3
    # the winner's cardinality agrees with an independent numpy fp64 count
    cnt, _, _, _ = rn.score(g.R_hyp, g.T_hyp, c.Ya, c.Yb, g.thr)
    assert cnt == g.best_fit


def test_hypothesis_block_split_equals_single_run(ctx, synth, pre3):
    """Config-5 sharding emulated on one GPU: the blocks' keys max-reduced give the same winner
    (max count, lowest id) as one run over all hypotheses, and finish() reproduces its result."""
    import torch
    c = synth.make_correspondences(31, N=5000, outlier_ratio=0.6)
    H, G = 20000, 4
    opts = pre3.make_opts(method=0, k=5, max_iteration=H + 1, adaptive=False, H=H, seed=17)
    g = ctx.ransac(c.Ya, c.Yb, None, opts)
    first = int(np.flatnonzero(g.counts == g.counts.max())[0])
    Ya = torch.from_numpy(c.Ya).cuda()
    Yb = torch.from_numpy(c.Yb).cuda()
    keys = []
    for r in range(G):
        key = torch.zeros(1, dtype=torch.int64, device="cuda")
        es = torch.zeros(1, dtype=torch.float64, device="cuda")
        ctx.ransac_block_dev(Ya, Yb, opts, r * (H // G), H // G, g.thr, key, es)
        ctx.sync()
        keys.append(int(key.item()))
    best = max(keys)
    assert best >> 32 == g.counts.max() and 0xFFFFFFFF - (best & 0xFFFFFFFF) == first
    res = torch.zeros(240, dtype=torch.uint8, device="cuda")
    mask = torch.zeros(5000, dtype=torch.uint8, device="cuda")
    ctx.ransac_finish_dev(Ya, Yb, opts, first, g.thr, res, mask)
    ctx.sync()
    r = pre3.unpack_result(np.frombuffer(res.cpu().numpy().tobytes(), dtype=pre3.RESULT_DTYPE)[0])
    assert r.best_fit == g.counts.max() and r.best_sample == first
    if first == g.best_sample:  # same winner as the (count, ErrorSum, first) rule -> same outputs
        np.testing.assert_array_equal(mask.cpu().numpy().astype(bool), g.mask)
        np.testing.assert_array_equal(r.R_hyp, g.R_hyp)  # the winning hypothesis itself: bit-exact
        np.testing.assert_array_equal(r.T_hyp, g.T_hyp)
        assert r.error_sum == g.error_sum
        # refit sums are reduced block-wide here and warp-wide in the batch path: rounding only
        np.testing.assert_allclose(r.R, g.R, rtol=0, atol=1e-13)
        np.testing.assert_allclose(r.T, g.T, rtol=0, atol=1e-13)


@pytest.mark.parametrize("supplied", [False, True])
def test_hypothesis_block_split_reference_mode(ctx, orc, synth, pre3, supplied):
    """Reference-exact mode of the split: every block's own (max count, min ErrorSum, first id) winner,
    16 bytes per block exchanged, same pick everywhere == the single run; the owning block already
    holds the mask and refit.  Also through dist.ransac_hypothesis_split without a process group."""
    import torch
    pd = importlib.import_module("3pre_b200.dist")
    N, H, G = 600, 4000, 4   # small N -> many hypotheses tie on the cardinality
    c = synth.make_correspondences(33, N=N, outlier_ratio=0.5, noise=0.0005)
    opts = pre3.make_opts(method=0, k=5, max_iteration=H + 1, adaptive=False, H=H, seed=23)
    samples = orc.sample_sets(23, 0, H, N, 5) if supplied else None
    g = ctx.ransac(c.Ya, c.Yb, samples, opts)
    Ya, Yb = torch.from_numpy(c.Ya).cuda(), torch.from_numpy(c.Yb).cuda()
    ds = torch.from_numpy(samples).cuda() if supplied else None
    counts, ids, ess, recs, masks = [], [], [], [], []
    for r in range(G):
        h0, h1 = pd.split_range(H, r, G)
        res = torch.zeros(240, dtype=torch.uint8, device="cuda")
        mask = torch.zeros(N, dtype=torch.uint8, device="cuda")
        ctx.ransac_block_select_dev(Ya, Yb, opts, h0, h1 - h0, g.thr, res, mask,
                                    samples=ds[h0:h1].contiguous() if supplied else None)
        ctx.sync()
        rec = np.frombuffer(res.cpu().numpy().tobytes(), dtype=pre3.RESULT_DTYPE)[0]
        assert rec["status"] == 0
        counts.append(int(rec["best_fit"])); ids.append(h0 + int(rec["best_sample"])); ess.append(float(rec["error_sum"]))
        recs.append(rec); masks.append(mask.cpu().numpy().astype(bool))
    w = pd.pick_reference(counts, ids, ess)
    assert (counts[w], ids[w], ess[w]) == (g.best_fit, g.best_sample, g.error_sum)
    np.testing.assert_array_equal(masks[w], g.mask)
    r = pre3.unpack_result(recs[w])
    np.testing.assert_array_equal(r.R, g.R)
    np.testing.assert_array_equal(r.T, g.T)
    assert r.state == g.state
    # the driver with world size 1 (no process group): both modes
    rec, m = pd.ransac_hypothesis_split(ctx, Ya, Yb, opts, samples=ds, mode="reference")
    assert (rec["best_fit"], rec["best_sample"], rec["error_sum"]) == (g.best_fit, g.best_sample, g.error_sum)
    np.testing.assert_array_equal(m.astype(bool), g.mask)
    rec, m = pd.ransac_hypothesis_split(ctx, Ya, Yb, opts, samples=ds, mode="first")
    assert rec["best_fit"] == g.counts.max() and rec["best_sample"] == int(np.flatnonzero(g.counts == g.counts.max())[0])


@pytest.mark.parametrize("mode", ["first", "reference"])
def test_split_stream_ordered_emulated(ctx, synth, pre3, mode):
    """The stream-ordered split (pre3_ransac_split_local_dev / _finish_dev), G ranks emulated one after the other on one
    GPU with the collective done by torch on the device: the summed records equal the single run; ranks that do not own
    the winner contribute zeros."""
    import torch
    pd = importlib.import_module("3pre_b200.dist")
    N, H, G = 600, 4000, 4
    c = synth.make_correspondences(33, N=N, outlier_ratio=0.5, noise=0.0005)
    opts = pre3.make_opts(method=0, k=5, max_iteration=H + 1, adaptive=False, H=H, seed=23)
    g = ctx.ransac(c.Ya, c.Yb, None, opts)
    Ya, Yb = torch.from_numpy(c.Ya).cuda(), torch.from_numpy(c.Yb).cuda()
    m = 0 if mode == "first" else 1
    torch.cuda.synchronize()
    keys, recs, masks = [], [], []
    for r in range(G):
        h0, h1 = pd.split_range(H, r, G)
        key = torch.zeros(2, dtype=torch.int64, device="cuda")
        res = torch.zeros(240, dtype=torch.uint8, device="cuda")
        mask = torch.zeros(N, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ctx.ransac_split_local_dev(Ya, Yb, opts, h0, h1 - h0, m, key, res if m else None, mask if m else None)
        ctx.sync()
        keys.append(key); recs.append(res); masks.append(mask)
    if m == 0:
        exchanged = torch.stack([k[0] for k in keys]).max().reshape(1).contiguous()      # all_reduce(MAX)
    else:
        exchanged = torch.cat(keys).contiguous()                                           # all_gather
    torch.cuda.synchronize()
    for r in range(G):
        h0, h1 = pd.split_range(H, r, G)
        ctx.ransac_split_finish_dev(Ya, Yb, opts, h0, h1 - h0, m, exchanged, G, r, recs[r], masks[r])
    ctx.sync()
    nonzero = [r for r in range(G) if recs[r].any().item()]
    assert len(nonzero) == 1                                                               # one owner
    total = torch.stack(recs).sum(0).to(torch.uint8)                                       # all_reduce(SUM)
    rec = np.frombuffer(total.cpu().numpy().tobytes(), dtype=pre3.RESULT_DTYPE)[0]
    mk = torch.stack(masks).sum(0).cpu().numpy().astype(bool)
    first = int(np.flatnonzero(g.counts == g.counts.max())[0])
    if mode == "first":
        assert (rec["best_fit"], rec["best_sample"]) == (g.counts.max(), first)
        h0o, h1o = pd.split_range(H, nonzero[0], G)
        assert h0o <= first < h1o
    else:
        assert (rec["best_fit"], rec["best_sample"], rec["error_sum"]) == (g.best_fit, g.best_sample, g.error_sum)
        np.testing.assert_array_equal(mk, g.mask)
        r_ = pre3.unpack_result(rec)
        np.testing.assert_array_equal(r_.R, g.R)
        np.testing.assert_array_equal(r_.T, g.T)
    assert mk.sum() == rec["best_fit"]


def test_matlab_mirror_roundtrip(ctx, orc, synth):
    m = importlib.import_module("3pre_b200.matlab")
    c = synth.make_correspondences(8, N=200, outlier_ratio=0.3)
    samples = synth.make_samples(9, 800, 200, 5)
    R, T, err, best_fit, state = m.RANSAC_CALC_VER2(c.Ya.T, c.Yb.T, {"DistanceThreshold": 0.05, "MaxIteration": 2000},
                                                    samples=samples.T + 1)
    o = orc.ransac(c.Ya, c.Yb, samples, method=0, max_iteration=2000)
    assert best_fit == o.best_fit and state == o.state and T.shape == (3, 1)
    assert err["mYa"].shape == (3, o.best_fit) and err["ErrorSum"] == o.error_sum
    assert rn.rot_angle(R, o.R) < TOL_ROT
    fp = synth.make_frame_pair(3, K1=300, K2=280, n_corr=150)
    matches, D = m.siftmatch(fp.desc1.T, fp.desc2.T, nargout=2)
    op, os_ = orc.siftmatch(fp.desc1, fp.desc2, 1.5)
    np.testing.assert_array_equal(matches, (op.T + 1).astype(float))
    np.testing.assert_array_equal(D, os_)
    s, Rh, Th, e = m.absoluteOrientationQuaternion(c.Yb[:50].T, c.Ya[:50].T)  # doScale defaults to 1
    s0, R0, T0, e0 = orc.horn(c.Yb[:50], c.Ya[:50], 1)
    assert s == s0 and e == e0
    Tq, q, Rr, st = m.Calculate_V_Omega_RANSAC_my_version(
        {"Descriptor": fp.desc1.T, "XYZ_DATA": fp.xyz1.T}, {"Descriptor": fp.desc2.T, "XYZ_DATA": fp.xyz2.T})
    assert q.shape == (4, 1) and abs(np.linalg.norm(q) - 1) < 1e-9 and st in (1, 2)
    assert rn.rot_angle(Rr, fp.R) < 2e-3


# ------------------------------------------------------------------------------------------
# the MEX gateway itself: 3pre_b200/mex_files/siftmatch.cpp built against the same stub mex
# runtime the reference's siftmatch.c is driven through (oracle/mex_stub), linked to libpre3.so
# ------------------------------------------------------------------------------------------
def _build_gateway():
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libpre3_siftmatch_gw.so")
    libdir = os.path.join(root, "3pre_b200", "lib")
    subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-o", so,
                    os.path.join(root, "3pre_b200", "mex_files", "siftmatch.cpp"),
                    "-x", "c", os.path.join(root, "oracle", "mex_stub", "mex_stub.c"), "-x", "none",
                    "-I", os.path.join(root, "oracle", "mex_stub"), "-I", os.path.join(root, "3pre_b200", "mex_files"),
                    "-I", os.path.join(root, "include"), "-L", libdir, "-lpre3", f"-Wl,-rpath,{libdir}"], check=True)
    return so


def test_mex_gateway_is_a_drop_in(golden, orc, synth):
    from oracle import refmex
    so = _build_gateway()
    for name in ["f64", "f32", "u8", "i8", "k2one", "ties", "knn"]:
        m, D = refmex.siftmatch(golden[f"{name}_L1"], golden[f"{name}_L2"], float(golden[f"{name}_thresh"]), nout=2, so=so)
        np.testing.assert_array_equal(m, golden[f"{name}_matches"])
        np.testing.assert_array_equal(D, golden[f"{name}_D"])
    fp = synth.make_frame_pair(123, K1=512, K2=512, n_corr=300)
    m, D = refmex.siftmatch(fp.desc1, fp.desc2, nout=2, so=so)  # default thresh 1.5
    if refmex.available():
        mr, Dr = refmex.siftmatch(fp.desc1, fp.desc2, nout=2)
        np.testing.assert_array_equal(m, mr)
        np.testing.assert_array_equal(D, Dr)
    op, os_ = orc.siftmatch(fp.desc1, fp.desc2, 1.5)
    np.testing.assert_array_equal(m, (op.T + 1).astype(float))
    # the reference's argument errors, raised through mexErrMsgTxt
    a = np.zeros((3, 4))
    with pytest.raises(refmex.MexError, match="same number of rows"):
        refmex.siftmatch(a, np.zeros((3, 5)), so=so)
    with pytest.raises(refmex.MexError, match="same class"):
        refmex.siftmatch(a, a.astype(np.float32), so=so)
    with pytest.raises(refmex.MexError, match="Unsupported numeric class"):
        refmex.siftmatch(a.astype(np.int32), a.astype(np.int32), so=so)
    with pytest.raises(refmex.MexError, match="At most three"):
        refmex.siftmatch(a, a, 1.5, extra_args=1, so=so)


def test_process_sequence_from_cached_sift_results(ctx, pre3, orc, synth, tmp_path):
    """The loop of find_consistent_sift_matches.m:22-32 over a folder of SIFT_result%04d.mat files (3pre_b200/formats.py):
    ragged feature counts per frame, one pre3_sequence call, RANSAC5_step files written as Calculate_V_Omega_RANSAC_my_version
    reads them."""
    import importlib
    fm = importlib.import_module("3pre_b200.formats")
    folder = str(tmp_path) + "/"
    frames = []
    base = synth.make_frame_pair(7700, K1=200, K2=200, n_corr=120)
    perm = np.random.default_rng(5).permutation(200)[:150]      # frame 2 re-observes frame 0's features, moved rigidly
    seq = [(base.desc1, base.xyz1), (base.desc2[:180], base.xyz2[:180]),
           (base.desc1[perm], base.xyz1[perm] @ base.R + np.array([0.02, -0.01, 0.03]))]
    for k, (d, x) in enumerate(seq):
        fm.save_sift_result(fm.sift_result_path(folder, 10 + k), {"idxScan": 10 + k, "Descriptor": d.T, "XYZ_DATA": x.T})
    o = pre3.make_opts(max_iteration=300, H=300, seed=4)
    res, matches, masks = fm.process_sequence(ctx, folder, 10, 12, opts=o)
    for p in range(2):
        om, g = orc.pair(seq[p][0], seq[p + 1][0], seq[p][1], seq[p + 1][1], 4, p, H=300, max_iteration=300)
        n = int(res["n_matches"][p])
        np.testing.assert_array_equal(matches[p, :n], om)
        assert (res["status"][p], res["best_fit"][p], res["best_sample"][p], res["state"][p]) == \
            (g.status, g.best_fit, g.best_sample, g.state), (p, res[p], g)
        T, R, st = fm.load_ransac_step(fm.ransac_step_path(folder, 10 + p, 11 + p))
        assert st == g.state and np.abs(R - g.R).max() < 1e-9 and np.abs(T.ravel() - g.T).max() < 1e-9


def test_sample_index_out_of_range_is_an_error(ctx, synth, pre3):
    """MATLAB raises 'Index exceeds matrix dimensions' for Ya(:, idx) with idx > N; explicit sample sets handed to the
    host entry points are validated the same way instead of being clamped."""
    c = synth.make_correspondences(5, N=50, outlier_ratio=0.2)
    samples = synth.make_samples(6, 20, 50, 5)
    bad = samples.copy(); bad[7, 2] = 50
    with pytest.raises(pre3.Pre3Error, match="Index exceeds matrix dimensions"):
        ctx.ransac(c.Ya, c.Yb, bad, pre3.make_opts(adaptive=False))
    neg = samples.copy(); neg[0, 0] = -1
    with pytest.raises(pre3.Pre3Error, match="Index exceeds matrix dimensions"):
        ctx.fit_batch(c.Ya, c.Yb, neg, 0)
    g = ctx.ransac(c.Ya, c.Yb, samples, pre3.make_opts(adaptive=False))   # the context stays usable
    assert g.status == 0

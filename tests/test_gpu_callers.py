"""GPU parity of the callers of siftmatch (-m gpu): the one-against-many sweep of
find_consistent_sift_matches.m:39-65 and matching_sift_based.m:104-150 (siftmatch + search-region gate), both
engines, all four descriptor classes, against the oracle on the same inputs."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _descs(synth, seed, K1, K2, P, cls):
    L1 = None
    L2 = []
    for p in range(P):
        fp = synth.make_frame_pair(seed + p, K1=K1, K2=K2, n_corr=min(K1, K2) // 2)
        if L1 is None:
            L1 = fp.desc1
            L2.append(fp.desc2)
        else:                       # later frames: noisy copies of part of the FIRST frame's descriptors
            rng = np.random.default_rng(seed + 100 + p)
            d = fp.desc2.copy()
            n = min(K1, K2) // 2
            src = rng.permutation(K1)[:n]
            dst = rng.permutation(K2)[:n]
            d[dst] = L1[src] + 0.02 * rng.standard_normal((n, 128)) * np.abs(L1[src]).max()
            L2.append(d)
    L2 = np.stack(L2)
    if cls == "u8":
        return synth.to_uint8(L1), synth.to_uint8(L2.reshape(-1, 128)).reshape(L2.shape)
    if cls == "i8":
        return (synth.to_uint8(L1) // 2).astype(np.int8), (synth.to_uint8(L2.reshape(-1, 128)) // 2).astype(np.int8).reshape(L2.shape)
    if cls == "f32":
        return L1.astype(np.float32), L2.astype(np.float32)
    return L1, L2


@pytest.mark.parametrize("cls", ["f64", "f32", "u8", "i8"])
@pytest.mark.parametrize("engine", [0, 1])
def test_sweep_vs_oracle(ctx, orc, synth, cls, engine):
    P, K1, K2 = 5, 300, 420
    L1, L2 = _descs(synth, 4000, K1, K2, P, cls)
    k2c = np.array([420, 400, 1, 0, 333], np.int32)
    ctx.set_match_engine(engine)
    try:
        out = ctx.siftmatch_sweep(L1, L2, 1.5, k2_count=k2c)
    finally:
        ctx.set_match_engine(0)
    total = 0
    for p in range(P):
        ref_pairs, ref_D = orc.siftmatch(L1, L2[p, : k2c[p]], 1.5)
        np.testing.assert_array_equal(out[p][0], ref_pairs)
        np.testing.assert_array_equal(out[p][1], ref_D)
        total += len(ref_pairs)
    assert total > 100


def test_sweep_equals_batch_with_repeated_L1(ctx, synth):
    P, K1, K2 = 9, 1024, 1100
    L1, L2 = _descs(synth, 4100, K1, K2, P, "u8")
    a = ctx.siftmatch_sweep(L1, L2, 1.5)
    b = ctx.siftmatch_batch(np.broadcast_to(L1, (P, K1, 128)).copy(), L2, 1.5)
    for p in range(P):
        np.testing.assert_array_equal(a[p][0], b[p][0])
        np.testing.assert_array_equal(a[p][1], b[p][1])
    # the batch entry right after the sweep is unaffected (the shared-L1 switch is scoped to the sweep call)
    c = ctx.siftmatch_batch(L2[:, :K1].copy(), L2, 1.5)
    assert all(len(c[p][0]) > 0 for p in range(P))


def test_sweep_dev_and_empty(ctx, synth):
    import torch
    P, K1, K2 = 3, 200, 256
    L1, L2 = _descs(synth, 4200, K1, K2, P, "f64")
    host = ctx.siftmatch_sweep(L1, L2, 1.5)
    d1, d2 = torch.from_numpy(L1).cuda(), torch.from_numpy(L2).cuda()
    pairs = torch.zeros((P, K1, 2), dtype=torch.int32, device="cuda")
    score = torch.zeros((P, K1), dtype=torch.float64, device="cuda")
    n = torch.zeros(P, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.siftmatch_sweep_dev(d1, d2, pairs, score, n)
    ctx.sync()
    for p in range(P):
        k = int(n[p])
        np.testing.assert_array_equal(pairs[p, :k].cpu().numpy(), host[p][0])
        np.testing.assert_array_equal(score[p, :k].cpu().numpy(), host[p][1])
    assert ctx.siftmatch_sweep(L1, L2[:0], 1.5) == []
    e = ctx.siftmatch_sweep(L1[:0], L2, 1.5)
    assert all(len(x[0]) == 0 for x in e)


@pytest.mark.parametrize("cls", ["f64", "u8"])
@pytest.mark.parametrize("engine", [0, 1])
def test_matching_sift_based_vs_oracle(ctx, orc, synth, cls, engine):
    P, F, K2 = 4, 150, 500
    rng = np.random.default_rng(77)
    des1 = np.zeros((P, F, 128)); des2 = np.zeros((P, K2, 128))
    for p in range(P):
        fp = synth.make_frame_pair(4300 + p, K1=F, K2=K2, n_corr=100)
        des1[p], des2[p] = fp.desc1, fp.desc2
    if cls == "u8":
        des1 = synth.to_uint8(des1.reshape(-1, 128)).reshape(P, F, 128)
        des2 = synth.to_uint8(des2.reshape(-1, 128)).reshape(P, K2, 128)
    pos2 = rng.uniform(0, 176, (P, K2, 2))
    fc = np.array([150, 120, 0, 150], np.int32)
    k2c = np.array([500, 500, 500, 3], np.int32)
    h = rng.uniform(0, 176, (P, F, 2))
    S11 = rng.uniform(1.0, 200.0, (P, F))
    S11[:, ::7] = np.nan                     # empty S: radius 40
    for p in range(P):                       # put most predictions near the matched raw feature, some exactly on the radius
        pr, _ = orc.siftmatch(des1[p, : fc[p]], des2[p, : k2c[p]], 1.5)
        for i, (k1, k2) in enumerate(pr):
            if i % 3 == 0:
                continue                     # stays random: usually outside the gate
            h[p, k1] = pos2[p, k2] + rng.uniform(-6, 6, 2)
            if i % 11 == 1:                  # dist == radius exactly (<= keeps it)
                rad = 40.0 if np.isnan(S11[p, i]) else np.ceil(3 * np.sqrt(S11[p, i]))
                h[p, k1] = pos2[p, k2] - np.array([rad, 0.0])
                pos2[p, k2] = h[p, k1] + np.array([rad, 0.0])
    ctx.set_match_engine(engine)
    try:
        r = ctx.matching_sift_based_batch(des1, des2, h, S11, pos2, f_count=fc, k2_count=k2c)
    finally:
        ctx.set_match_engine(0)
    seen_in = seen_out = 0
    for p in range(P):
        f = fc[p]
        ic, z, mt, nm, nd = orc.matching_sift_based(des1[p, :f], des2[p, : k2c[p]], h[p, :f], S11[p, :f], pos2[p, : k2c[p]])
        np.testing.assert_array_equal(r["ic"][p, :f], ic)
        np.testing.assert_array_equal(r["match"][p, :f], mt)
        np.testing.assert_array_equal(r["z"][p, :f], z)   # NaN == NaN in assert_array_equal
        assert not r["ic"][p, f:].any() and (r["match"][p, f:] == -1).all()
        assert r["n_match"][p] == nm and r["n_discarded"][p] == nd
        seen_in += int(ic.sum()); seen_out += nd
    assert seen_in > 50 and seen_out > 20


def test_matlab_matching_sift_based_and_sweep(orc, synth):
    M = importlib.import_module("3pre_b200.matlab")
    fp = synth.make_frame_pair(4400, K1=80, K2=300, n_corr=60)
    des2 = synth.to_uint8(fp.desc2).T.copy()          # ND x K2 (uint8, as SIFT_extract_save stores it)
    d1 = synth.to_uint8(fp.desc1)
    rng = np.random.default_rng(5)
    pos = np.vstack([rng.uniform(0, 176, (2, 300)), rng.uniform(0, 5, (2, 300))])
    pr, _ = orc.siftmatch(d1, des2.T.copy(), 1.5)
    feats = []
    for i in range(100):
        if i % 5 == 4:
            feats.append({"h": np.zeros((0, 0)), "S": np.zeros((0, 0)), "Descriptor": np.zeros(128, np.uint8),
                          "individually_compatible": 0, "z": np.zeros((0, 0))})
            continue
        j = len([f for f in feats if np.size(f["h"])])
        feats.append({"h": rng.uniform(0, 176, (1, 2)), "S": np.zeros((0, 0)) if i % 3 == 0 else np.diag([30.0, 30.0]),
                      "Descriptor": d1[j].astype(np.float64).reshape(128, 1), "individually_compatible": 0,
                      "z": np.zeros((0, 0))})
    idx = [i for i, f in enumerate(feats) if np.size(f["h"])]
    for k1, k2 in pr[::2]:
        feats[idx[k1]]["h"] = (pos[:2, k2] + 1.5).reshape(1, 2)
    scan = {"Descriptor_RAW": des2, "SCALE_ORIENT_POS_RAW": pos}
    ref_ic, ref_z, ref_mt, nm, nd = orc.matching_sift_based(
        d1[: len(idx)], des2.T.copy(), np.stack([feats[i]["h"].reshape(-1) for i in idx]),
        np.array([np.nan if not np.size(feats[i]["S"]) else feats[i]["S"][0, 0] for i in idx]), pos[:2].T.copy())
    out, disc = M.matching_sift_based(scan, feats, step_global=12)
    assert disc == nd and ref_ic.sum() >= 10
    for j, i in enumerate(idx):
        assert bool(out[i]["individually_compatible"]) == bool(ref_ic[j])
        if ref_ic[j]:
            np.testing.assert_array_equal(out[i]["z"], ref_z[j])
            assert out[i]["last_visible"] == 12
            np.testing.assert_array_equal(out[i]["Descriptor"], des2[:, ref_mt[j]])
    # nothing predicted: returned untouched (:114-116)
    none, d0 = M.matching_sift_based(scan, [{"h": np.zeros((0, 0))}], 0)
    assert d0 == 0 and list(none[0]) == ["h"]
    sw = M.siftmatch_sweep(d1.T, [des2, des2[:, :100], des2[:, :0]])
    for m, d in zip(sw, [des2, des2[:, :100], des2[:, :0]]):
        ref, _ = orc.siftmatch(d1, d.T.copy(), 1.5)
        np.testing.assert_array_equal(m, (ref.T + 1).astype(np.float64))
    with pytest.raises(M.MexError, match="same class"):
        M.siftmatch_sweep(d1.T, [des2.astype(np.float64)])

"""GPU parity of the integer tensor-core matcher (tcgen05 kind::i8, csrc/match_i8.cu; -m gpu): class int8 / uint8 against
the reference's own compiled siftmatch.c (oracle/_ref) or the oracle, at the extremes of the value range (signedness of
both operands, largest distances), with ties, duplicate columns, ragged batches, and through the forced engine."""
import numpy as np
import pytest

from oracle import refmex

pytestmark = pytest.mark.gpu


def _ref(orc, L1, L2, thr):
    if refmex.available():
        m, D = refmex.siftmatch(L1, L2, thr, nout=2)
        return (m.T - 1).astype(np.int32).reshape(-1, 2), np.asarray(D).reshape(-1)
    return orc.siftmatch(L1, L2, thr)


@pytest.mark.parametrize("dtype", [np.uint8, np.int8])
@pytest.mark.parametrize("engine", [0, 2])
def test_full_range_values(ctx, orc, dtype, engine):
    rng = np.random.default_rng(11)
    lo, hi = (0, 255) if dtype == np.uint8 else (-128, 127)
    K1, K2 = 300, 700
    L1 = rng.integers(lo, hi + 1, (K1, 128)).astype(dtype)
    L2 = rng.integers(lo, hi + 1, (K2, 128)).astype(dtype)
    L1[0], L1[1], L1[2] = lo, hi, lo          # constant rows at the extremes
    L2[0], L2[1] = hi, lo                     # the largest possible distance 128 * 255^2 = 8 323 200
    L2[5:205] = L1[40:240]                    # exact hits (distance 0, accepted whatever the runner-up)
    L2[300] = L2[6]                           # duplicate of a hit: best == second == 0 -> 1.5 * 0 <= 0 accepted, first index
    L2[400:420] = np.clip(L1[250:270].astype(np.int32) + rng.integers(-3, 4, (20, 128)), lo, hi).astype(dtype)
    ctx.set_match_engine(engine)
    try:
        for thr in (1.5, 1.0, 4.0):
            pairs, score = ctx.siftmatch(L1, L2, thr)
            rp, rs = _ref(orc, L1, L2, thr)
            np.testing.assert_array_equal(pairs, rp)
            np.testing.assert_array_equal(score, rs)
        assert len(rp) >= 200
        # one column only: second_best stays at INT_MAX (siftmatch.c:66-67), every row accepted
        pairs, score = ctx.siftmatch(L1, L2[:1], 1.5)
        rp, rs = _ref(orc, L1, L2[:1], 1.5)
        np.testing.assert_array_equal(pairs, rp)
        np.testing.assert_array_equal(score, rs)
        assert len(rp) == K1 and score.max() == 128.0 * 255 * 255
    finally:
        ctx.set_match_engine(0)


@pytest.mark.parametrize("dtype", [np.uint8, np.int8])
def test_ties_across_tiles_first_index_wins(ctx, orc, dtype):
    """Equal distances in different 128-column tiles and in different halves of the CTA pair: the first index wins."""
    rng = np.random.default_rng(12)
    K1, K2 = 260, 1200
    L1 = rng.integers(0, 100, (K1, 128)).astype(dtype)
    L2 = rng.integers(0, 100, (K2, 128)).astype(dtype)
    for i, cols in enumerate([(900, 130), (64, 63), (1199, 0), (127, 128), (500, 1000, 250)]):
        v = L1[i * 50].copy()
        v[0] += 1                                   # distance 1 to L1 row, in several columns
        for c in cols:
            L2[c] = v
    pairs, score = ctx.siftmatch(L1, L2, 1.0)       # thresh 1: best <= second accepted
    rp, rs = _ref(orc, L1, L2, 1.0)
    np.testing.assert_array_equal(pairs, rp)
    np.testing.assert_array_equal(score, rs)
    got = {int(a): int(b) for a, b in pairs}
    assert got[0] == 130 and got[50] == 63 and got[100] == 0 and got[150] == 127 and got[200] == 250


def test_batch_ragged_and_misaligned_fallback(ctx, orc, synth):
    import torch
    P, K1, K2 = 6, 520, 300
    L1 = np.zeros((P, K1, 128), np.uint8); L2 = np.zeros((P, K2, 128), np.uint8)
    for p in range(P):
        fp = synth.make_frame_pair(900 + p, K1=K1, K2=K2, n_corr=200)
        L1[p], L2[p] = synth.to_uint8(fp.desc1), synth.to_uint8(fp.desc2)
    k1c = np.array([520, 513, 0, 1, 256, 257], np.int32)
    k2c = np.array([300, 1, 128, 0, 129, 300], np.int32)
    out = ctx.siftmatch_batch(L1, L2, 1.5, k1c, k2c)
    for p in range(P):
        rp, rs = orc.siftmatch(L1[p, : k1c[p]], L2[p, : k2c[p]], 1.5)
        np.testing.assert_array_equal(out[p][0], rp)
        np.testing.assert_array_equal(out[p][1], rs)
    # a descriptor set that is not 16-byte aligned goes to the exact kernel (same results)
    raw = torch.zeros(P * K1 * 128 + 16, dtype=torch.uint8, device="cuda")
    raw[1: 1 + P * K1 * 128] = torch.from_numpy(L1.reshape(-1)).cuda()
    d1 = raw[1: 1 + P * K1 * 128].view(P, K1, 128)
    d2 = torch.from_numpy(L2).cuda()
    pairs = torch.zeros((P, K1, 2), dtype=torch.int32, device="cuda")
    score = torch.zeros((P, K1), dtype=torch.float64, device="cuda")
    n = torch.zeros(P, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.siftmatch_batch_dev(d1, d2, pairs, score, n, 1.5, torch.from_numpy(k1c).cuda(), torch.from_numpy(k2c).cuda())
    ctx.sync()
    for p in range(P):
        k = int(n[p])
        np.testing.assert_array_equal(pairs[p, :k].cpu().numpy(), out[p][0])
        np.testing.assert_array_equal(score[p, :k].cpu().numpy(), out[p][1])


def test_cfg2_uint8_through_the_integer_tensor_cores(ctx, orc, synth):
    """BASELINE cfg2's uint8 variant (2048 x 2048 x 128, round(512 d)): timing spans show the GEMM engine ran, and the
    exact engine gives the same rows."""
    P, K = 3, 2048
    L1 = np.zeros((P, K, 128), np.uint8); L2 = np.zeros((P, K, 128), np.uint8)
    for p in range(P):
        fp = synth.make_frame_pair(3000 + p, K1=K, K2=K, n_corr=K // 2)
        L1[p], L2[p] = synth.to_uint8(fp.desc1), synth.to_uint8(fp.desc2)
    ctx.timing_enable(True)
    try:
        ctx.timing_read()
        a = ctx.siftmatch_batch(L1, L2, 1.5)
        t = ctx.timing_read()
    finally:
        ctx.timing_enable(False)
    assert t["match_tc"][1] == 1 and "match_exact" not in t and "rescore" not in t
    ctx.set_match_engine(1)
    try:
        b = ctx.siftmatch_batch(L1, L2, 1.5)
    finally:
        ctx.set_match_engine(0)
    for p in range(P):
        np.testing.assert_array_equal(a[p][0], b[p][0])
        np.testing.assert_array_equal(a[p][1], b[p][1])
        assert len(a[p][0]) >= 900

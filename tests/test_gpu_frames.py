"""GPU parity: SR4000 frame batches -> filtered maps -> per-feature 3-D points through the C ABI
vs the CPU oracle (oracle/pre3_oracle_frames.c; same tap order -> bit-exact values, keep flags and idxRemain) and the
committed golden vectors of the independent scipy restatement (1e-12)."""
import importlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "frames.npz"))


def _eq_nan(a, b):
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_array_equal(np.nan_to_num(a, nan=-7.0), np.nan_to_num(b, nan=-7.0))


@pytest.mark.parametrize("rows", [720, 576, 721])
@pytest.mark.parametrize("flavour", ["sift_extract_save", "dr_ye"])
def test_features_vs_oracle(ctx, orc, synth, rows, flavour):
    F, K = 5, 300
    sr, fr = synth.make_sr_frames(300 + rows, F, K, rows=rows)
    kw = dict(sigma=2.0, boundary=0, mode=0) if flavour == "sift_extract_save" else dict(sigma=1.0, boundary=1, mode=1)
    desc = np.random.default_rng(1).random((F, K, 128))
    kc = np.array([K, K - 7, 1, 0, K], np.int32)
    out = ctx.features_xyz_batch(sr, fr, desc=desc, k_count=kc, **kw)
    for f in range(F):
        n = int(kc[f])
        xyz, keep, idx, oob = orc.features_xyz(sr[f], fr[f, :n], use_conf=1, **kw)
        _eq_nan(out["xyz_all"][f, :n], xyz)
        assert np.isnan(out["xyz_all"][f, n:]).all() and not out["keep"][f, n:].any()
        np.testing.assert_array_equal(out["keep"][f, :n], keep)
        m = len(idx)
        assert out["n_keep"][f] == m
        np.testing.assert_array_equal(out["idx_remain"][f, :m], idx)
        assert (out["idx_remain"][f, m:] == -1).all()
        _eq_nan(out["xyz"][f, :m], xyz[idx])                                   # XYZ_DATA = xyz_data(:, idxRemain)
        np.testing.assert_array_equal(out["desc_out"][f, :m], desc[f, idx])     # Descriptor(:, idxRemain)
        np.testing.assert_array_equal(out["frames_out"][f, :m], fr[f, idx])     # SCALE_ORIENT_POS(:, idxRemain)
        assert not out["desc_out"][f, m:].any()
    assert out["n_oob"] == 0


def test_maps_vs_oracle_and_fused_lookup(ctx, orc, synth):
    F = 3
    sr, fr = synth.make_sr_frames(411, F, 64)
    for sigma, boundary in ((2.0, 0), (1.0, 1)):
        x, y, z, mc = ctx.read_xyz_sr4000_batch(sr, sigma=sigma, boundary=boundary)
        for f in range(F):
            ox, oy, oz = orc.read_xyz(sr[f], sigma, boundary)
            _eq_nan(x[f], ox); _eq_nan(y[f], oy); _eq_nan(z[f], oz)
            assert mc[f] == orc.max_confidence(sr[f])
    # 576-row frames have no confidence map
    _, _, _, mc = ctx.read_xyz_sr4000_batch(np.ascontiguousarray(sr[:, :, :576]))
    assert np.isnan(mc).all()


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_golden(ctx, gold, name):
    rows = {"a": 720, "b": 576, "c": 721}[name]
    sr = np.ascontiguousarray(gold["sr"].astype(np.float64)[:, :rows])[None]
    fr = gold["frames"].astype(np.float64)[None]
    out = ctx.features_xyz_batch(sr, fr)
    n = int(out["n_keep"][0])
    np.testing.assert_array_equal(out["idx_remain"][0, :n], gold[f"{name}_remain"])
    g = gold[f"{name}_xyz"].T
    np.testing.assert_array_equal(np.isnan(out["xyz_all"][0]), np.isnan(g))
    assert np.nanmax(np.abs(out["xyz_all"][0] - g)) < 1e-12
    x, y, z, _ = ctx.read_xyz_sr4000_batch(sr)
    assert np.nanmax(np.abs(z[0].T[70] - gold[f"{name}_zrow"])) < 1e-12
    assert np.nanmax(np.abs(x[0].T[:, 0] - gold[f"{name}_xcol"])) < 1e-12
    _, yr, _, _ = ctx.read_xyz_sr4000_batch(sr, sigma=1.0, boundary=1)
    assert np.nanmax(np.abs(yr[0].T[:, 175] - gold[f"{name}_ycol_dr_ye"])) < 1e-12


def test_out_of_image_and_empty(ctx, pre3, synth):
    sr, fr = synth.make_sr_frames(5, 2, 8)
    fr[0, 3, 0] = 175.6      # rounds to column 177
    fr[1, 0, 1] = -1.6       # rounds to row 0
    out = ctx.features_xyz_batch(sr, fr)
    assert out["n_oob"] == 2 and not out["keep"][0, 3] and not out["keep"][1, 0]
    out = ctx.features_xyz_batch(sr, fr[:, :0])
    assert (out["n_keep"][:2] == 0).all()
    with pytest.raises(pre3.Pre3Error):
        ctx.features_xyz_batch(np.zeros((1, 176, 700)), fr[:1])


def test_matlab_mirror(orc, synth):
    ml = importlib.import_module("3pre_b200.matlab")
    from oracle import ref_numpy as rn
    sr, fr = synth.make_sr_frames(77, 1, 120)
    M = sr[0].T.copy()                           # the 720 x 176 matrix `load` returns
    x, y, z, cm = ml.read_xyz_sr4000(M)
    gx, gy, gz, gcm = rn.read_xyz_sr4000(M)
    assert np.nanmax(np.abs(x - gx)) < 1e-12 and np.nanmax(np.abs(z - gz)) < 1e-12 and np.array_equal(cm, gcm)
    desc = np.random.default_rng(2).random((128, 120))
    S = ml.SIFT_extract_save(M, fr[0].T, desc, idxScan=12)
    g, rem = rn.sift_extract_xyz(M, fr[0].T)
    np.testing.assert_array_equal(S["idxRemain"], rem + 1)
    assert np.abs(S["XYZ_DATA"] - g[:, rem]).max() < 1e-12
    np.testing.assert_array_equal(S["Descriptor"], desc[:, rem])
    np.testing.assert_array_equal(S["SCALE_ORIENT_POS"][:2], fr[0].T[:2, rem] + 1)
    np.testing.assert_array_equal(S["SCALE_ORIENT_POS_RAW"][:2], fr[0].T[:2] + 1)
    k = int(rem[0])
    rho, p = ml.inittialize_depth_my_version(fr[0, k, :2] + 1, M)
    assert np.abs(p - g[:, k]).max() < 1e-12 and abs(rho - 1 / np.linalg.norm(p)) < 1e-15
    bad = int(np.setdiff1d(np.arange(120), rem)[0])
    assert ml.inittialize_depth_my_version(fr[0, bad, :2] + 1, M) == (None, None)
    xr, yr, zr, _ = ml.read_sr4000_data_dr_ye(M)
    gx, gy, gz, _ = rn.read_xyz_sr4000(M, 1.0, "replicate")
    assert np.nanmax(np.abs(yr - gy)) < 1e-12


def test_frames_feed_the_sequence_path(ctx, pre3, synth):
    """frames -> features_xyz (device) -> pre3_sequence_dev with k_count: the compacted descriptors / points of the
    step before the path are exactly what the matching path consumes."""
    torch = pytest.importorskip("torch")
    F, K = 4, 256
    sr, fr = synth.make_sr_frames(900, F, K)
    desc = synth.make_frame_pair(1, K1=K, K2=K, n_corr=100).desc1
    d = torch.tensor(np.broadcast_to(desc, (F, K, 128)).copy(), device="cuda")
    o = pre3.make_frame_opts()
    xyz = torch.zeros(F, K, 3, dtype=torch.float64, device="cuda")
    nk = torch.zeros(F, dtype=torch.int32, device="cuda")
    dout = torch.zeros_like(d)
    ctx.features_xyz_batch_dev(torch.tensor(sr, device="cuda"), o, torch.tensor(fr, device="cuda"), xyz=xyz, n_keep=nk,
                               desc_in=d, desc_out=dout)
    res = torch.zeros(F - 1, 240, dtype=torch.uint8, device="cuda")
    ctx.sequence_dev(dout, torch.nan_to_num(xyz), pre3.make_opts(H=64, max_iteration=64, seed=1), res, k_count=nk)
    ctx.sync()
    rec = np.frombuffer(res.cpu().numpy().tobytes(), dtype=pre3.RESULT_DTYPE)
    ref = ctx.features_xyz_batch(sr, fr)
    np.testing.assert_array_equal(nk.cpu().numpy(), ref["n_keep"])
    assert (rec["n_matches"] <= ref["n_keep"][:-1]).all() and (rec["n_matches"] > 0).all()

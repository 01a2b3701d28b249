"""CPU suite: the oracle of the frame -> per-feature 3-D point step (oracle/pre3_oracle_frames.c).

PARITY UNPINNED (no reference vectors; fspecial / imfilter absent): the C restatement is checked against the
independent scipy.ndimage restatement (oracle/ref_numpy.py; fixtures tests/golden/frames.npz made by
tests/golden/make_golden_frames.py) and on the rejection rules read off M/inittialize_depth_my_version.m.
"""
import importlib
import os

import numpy as np
import pytest

from oracle import ref_numpy as rn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "frames.npz"))


def _case(gold, name):
    rows = {"a": 720, "b": 576, "c": 721}[name]
    sr = np.ascontiguousarray(gold["sr"].astype(np.float64)[:, :rows])
    return sr, gold["frames"].astype(np.float64)


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_oracle_vs_golden(orc, gold, name):
    sr, fr = _case(gold, name)
    xyz, keep, idx, oob = orc.features_xyz(sr, fr)
    np.testing.assert_array_equal(idx, gold[f"{name}_remain"])          # idxRemain (SIFT_extract_save.m:82)
    np.testing.assert_array_equal(np.isnan(xyz.T), np.isnan(gold[f"{name}_xyz"]))
    assert np.nanmax(np.abs(xyz.T - gold[f"{name}_xyz"])) < 1e-12 and oob == 0
    x, y, z = orc.read_xyz(sr)
    g = gold[f"{name}_zrow"]
    np.testing.assert_array_equal(np.isnan(z.T[70]), np.isnan(g))
    assert np.nanmax(np.abs(z.T[70] - g)) < 1e-12
    g = gold[f"{name}_xcol"]
    assert np.nanmax(np.abs(x.T[:, 0] - g)) < 1e-12                     # zero padding at the border
    xr, yr, zr = orc.read_xyz(sr, sigma=1.0, boundary=1)
    assert np.nanmax(np.abs(yr.T[:, 175] - gold[f"{name}_ycol_dr_ye"])) < 1e-12   # 'replicate'
    # the fused per-feature stencil reads the same values as the full maps
    r = np.floor(fr[:, 1] + 1 + 0.5).astype(int) - 1
    c = np.floor(fr[:, 0] + 1 + 0.5).astype(int) - 1
    np.testing.assert_array_equal(xyz[keep, 0], -x[c[keep], r[keep]])
    np.testing.assert_array_equal(xyz[keep, 2], z[c[keep], r[keep]])


def test_kernel_is_fspecial(orc):
    for s in (1.0, 2.0, 0.5):
        assert np.abs(orc.gaussian3(s) - rn.fspecial_gaussian3(s)).max() < 1e-16
        assert abs(orc.gaussian3(s).sum() - 1) < 1e-15


def test_rejection_rules(orc):
    # a flat scene 2 m away, confidence 100 everywhere
    sr = np.zeros((176, 720))
    sr[:, 0:144] = 2.0
    sr[:, 576:720] = 100.0
    fr = np.array([[10.0, 10.0, 1, 0],      # fine
                   [50.0, 50.0, 1, 0],      # NaN in x nearby -> smoothed x is NaN -> rejected (:40)
                   [90.0, 90.0, 1, 0],      # confidence exactly half the maximum -> rejected (<=, :74)
                   [120.0, 30.0, 1, 0],     # closer than 0.4 m -> rejected
                   [20.49, 20.5, 1, 0],     # round half away from zero: (x+1, y+1) = (21.49, 21.5) -> column 21, row 22
                   [130.0, 100.0, 1, 0]])   # y map NaN only: x is not NaN, df = NaN, `df < 0.4` false -> KEPT with NaN
    sr[51, 144 + 50] = np.nan
    sr[90, 576 + 90] = 50.0
    sr[119:122, 29:32] = 0.1
    sr[130, 288 + 100] = np.nan
    sr[20, 21] = 7.0                         # z(row 22, column 21) in 1-based terms
    xyz, keep, idx, oob = orc.features_xyz(sr, fr)
    np.testing.assert_array_equal(keep, [True, False, False, False, True, True])
    np.testing.assert_array_equal(idx, [0, 4, 5])
    h = orc.gaussian3(2.0)
    assert abs(xyz[4, 2] - (2.0 + h[1, 1] * 5.0)) < 1e-15
    assert np.isnan(xyz[5, 1]) and abs(xyz[5, 2] - 2.0) < 1e-15
    assert np.isnan(xyz[1]).all()
    # zero padding: a corner pixel keeps only 4 of the 9 taps
    xyz, keep, _, _ = orc.features_xyz(sr, np.array([[0.0, 0.0, 1, 0]]))
    assert abs(xyz[0, 2] - 2.0 * h[1:, 1:].sum()) < 1e-15
    # without a confidence map (576 rows) only NaN / 0.4 m reject
    _, keep, _, _ = orc.features_xyz(np.ascontiguousarray(sr[:, :576]), fr)
    np.testing.assert_array_equal(keep, [True, False, True, False, True, True])
    # dr_ye flavour: confidence strictly below half the maximum removes the feature (confidence_filtering.m:8)
    _, keep, _, _ = orc.features_xyz(sr, fr, sigma=1.0, boundary=1, mode=1, use_conf=1)
    np.testing.assert_array_equal(keep, [True, True, True, True, True, True])
    sr[90, 576 + 90] = 49.0
    _, keep, _, _ = orc.features_xyz(sr, fr, sigma=1.0, boundary=1, mode=1, use_conf=1)
    np.testing.assert_array_equal(keep, [True, True, False, True, True, True])
    # outside the image: counted (the reference raises an index error)
    _, keep, _, oob = orc.features_xyz(sr, np.array([[175.6, 10.0, 1, 0], [10.0, -1.6, 1, 0]]))
    assert oob == 2 and not keep.any()

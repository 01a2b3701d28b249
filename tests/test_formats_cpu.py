"""CPU suite: the on-disk formats of the reference's caches (3pre_b200/formats.py): file names as
the reference builds them, exact round trips of d1_%04d.dat / SIFT_result%04d.mat / RANSAC5_step_%d_%d.mat."""
import importlib

import numpy as np

fm = importlib.import_module("3pre_b200.formats")
synth = importlib.import_module("3pre_b200.synth")


def test_file_names_follow_the_reference():
    assert fm.d1_path("/data/run1", 7) == "/data/run1/d1_0007.dat"                                   # read_xyz_sr4000.m:2
    assert fm.sift_result_path("/data/run1/", 12) == "/data/run1/FeatureExtractionMatching/SIFT_result0012.mat"
    assert fm.ransac_step_path("/data/run1/", 3, 4) == "/data/run1//RANSAC_pose_shift/RANSAC5_step_3_4.mat"


def test_d1_round_trip_is_exact(tmp_path):
    sr, _ = synth.make_sr_frames(5, 1, 4, rows=721, n_nan=50)
    M = sr[0].T.copy()                              # 721 x 176 as `load` returns it, NaNs included
    p = fm.d1_path(str(tmp_path), 3)
    fm.save_d1(p, M)
    back = fm.load_d1(p)
    assert back.shape == (721, 176)
    np.testing.assert_array_equal(np.isnan(back), np.isnan(M))
    np.testing.assert_array_equal(np.nan_to_num(back), np.nan_to_num(M))


def test_mat_round_trips(tmp_path):
    fp = synth.make_frame_pair(9, K1=40, K2=40, n_corr=20)
    S = {"idxScan": 17, "Image": np.zeros((144, 176), np.uint8), "Descriptor_RAW": fp.desc1.T, "SCALE_ORIENT_POS_RAW":
         np.ones((4, 40)), "Descriptor": fp.desc1.T[:, :30], "SCALE_ORIENT_POS": np.ones((4, 30)), "XYZ_DATA": fp.xyz1.T[:, :30]}
    p = fm.sift_result_path(str(tmp_path) + "/", 17)
    fm.save_sift_result(p, S)
    B = fm.load_sift_result(p)
    assert B["idxScan"] == 17 and B["Image"].dtype == np.uint8
    for k in ("Descriptor", "XYZ_DATA", "Descriptor_RAW"):
        np.testing.assert_array_equal(B[k], S[k])
    q = fm.ransac_step_path(str(tmp_path), 1, 2)
    fm.save_ransac_step(q, fp.R, fp.t, 1, best_fit=210, matches=np.array([[1, 2], [3, 4]]))
    T, R, st = fm.load_ransac_step(q)
    assert st == 1 and np.array_equal(R, fp.R) and np.array_equal(T.ravel(), fp.t)

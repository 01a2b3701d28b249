"""On-disk formats of the reference's caches, so that the drop-ins run on real SR4000 logs.

Host-side file I/O only (numpy / scipy.io) -- nothing here is on the GPU path:
  d1_%04d.dat                 one SR4000 frame: ASCII matrix, 576 / 720 / 721 rows x 176 columns, what
                              `load(sprintf('%s/d1_%04d.dat', prefix, k))` reads (M/read_xyz_sr4000.m:2-3)
  FeatureExtractionMatching/SIFT_result%04d.mat   variable SCAN_SIFT (M/SIFT_extract_save.m:44-45,68-69,86-88,103-106)
  RANSAC_pose_shift/RANSAC5_step_%d_%d.mat        the workspace of SIFT_match_save (M/SIFT_match_save.m:79-80), of which
                              M/Calculate_V_Omega_RANSAC_my_version.m:9 loads T_RANSAC, R_RANSAC, State_RANSAC
`process_sequence` is the loop of M/find_consistent_sift_matches.m:22-32 / RANSAC_CALC_SAVE_SR4000.m:14-15 over the
cached SIFT results of a folder: one pre3_sequence call instead of one MATLAB call per pair.
"""
from __future__ import annotations

import os

import numpy as np

SCAN_SIFT_FIELDS = ("idxScan", "Image", "Descriptor_RAW", "SCALE_ORIENT_POS_RAW", "Descriptor", "SCALE_ORIENT_POS",
                    "XYZ_DATA")


def d1_path(prefix: str, k: int) -> str:
    return "%s/d1_%04d.dat" % (prefix, k)                       # read_xyz_sr4000.m:2


def sift_result_path(data_folder: str, k: int) -> str:
    return "%sFeatureExtractionMatching/SIFT_result%04d.mat" % (data_folder, k)   # SIFT_extract_save.m:104


def ransac_step_path(data_folder: str, i: int, j: int) -> str:
    return "%s/RANSAC_pose_shift/RANSAC5_step_%d_%d.mat" % (data_folder, i, j)    # SIFT_match_save.m:79


def load_d1(path: str) -> np.ndarray:
    """sr_data = load(path): rows x 176 double (rows = 576, 720 or 721; a short last row -- the time stamp line of
    721-row files holds one number -- is padded with zeros like MATLAB would refuse to: such files are written by
    save_d1 with a full row)."""
    rows = []
    with open(path) as f:
        for line in f:
            v = np.array(line.split(), dtype=np.float64)
            if v.size:
                rows.append(v)
    width = max(r.size for r in rows)
    out = np.zeros((len(rows), width))
    for i, r in enumerate(rows):
        out[i, : r.size] = r
    return out


def save_d1(path: str, sr_data: np.ndarray) -> None:
    """ASCII writer for synthetic logs; %.17g keeps doubles exact through the text round trip."""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    np.savetxt(path, np.asarray(sr_data, np.float64), fmt="%.17g")


def save_sift_result(path: str, scan_sift: dict) -> None:
    from scipy.io import savemat
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    savemat(path, {"SCAN_SIFT": {k: scan_sift[k] for k in SCAN_SIFT_FIELDS if k in scan_sift}}, do_compression=True)


def load_sift_result(path: str) -> dict:
    """temp = load(path, 'SCAN_SIFT'); returns the struct as a dict of arrays (SIFT_match_save.m:20-21)."""
    from scipy.io import loadmat
    s = loadmat(path, squeeze_me=False, struct_as_record=False)["SCAN_SIFT"][0, 0]
    out = {}
    for k in SCAN_SIFT_FIELDS:
        if hasattr(s, k):
            v = getattr(s, k)
            out[k] = v.item() if k == "idxScan" and np.size(v) == 1 else np.asarray(v)
    return out


def save_ransac_step(path: str, R, T, state, best_fit=None, matches=None) -> None:
    from scipy.io import savemat
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    d = {"R_RANSAC": np.asarray(R, np.float64), "T_RANSAC": np.asarray(T, np.float64).reshape(3, 1),
         "State_RANSAC": float(state)}
    if best_fit is not None:
        d["BestFit"] = float(best_fit)
    if matches is not None:
        d["matches"] = np.asarray(matches, np.float64)
    savemat(path, d)


def load_ransac_step(path: str):
    """load(RANSAC_FileName, 'T_RANSAC', 'R_RANSAC', 'State_RANSAC') (Calculate_V_Omega_RANSAC_my_version.m:9)."""
    from scipy.io import loadmat
    d = loadmat(path)
    return d["T_RANSAC"].reshape(3, 1), d["R_RANSAC"], int(d["State_RANSAC"].ravel()[0])


def process_sequence(ctx, data_folder: str, first: int, last: int, opts=None, write: bool = True, K: int | None = None):
    """RANSAC_CALC_SAVE_SR4000(i, i+1) for i = first .. last-1 (M/find_consistent_sift_matches.m:22-32) from the cached
    SIFT_result files: frames are padded to a common feature count K and sent through ONE pre3_sequence call; the results
    are written as RANSAC5_step_%d_%d.mat.  Returns (results, matches, masks) of Context.sequence."""
    from .api import make_opts
    scans = [load_sift_result(sift_result_path(data_folder, k)) for k in range(first, last + 1)]
    counts = np.array([s["Descriptor"].shape[1] for s in scans], np.int32)
    K = int(K or max(1, counts.max()))
    F = len(scans)
    desc = np.zeros((F, K, 128))
    xyz = np.zeros((F, K, 3))
    for f, s in enumerate(scans):
        m = int(counts[f])
        desc[f, :m] = np.asarray(s["Descriptor"], np.float64).T
        xyz[f, :m] = np.asarray(s["XYZ_DATA"], np.float64).T
    o = opts or make_opts(max_iteration=2000, H=2000)            # SIFT_match_save.m:49-50
    res, matches, masks = ctx.sequence(desc, xyz, o, k_count=counts)
    if write:
        for p in range(F - 1):
            n = int(res["n_matches"][p])
            save_ransac_step(ransac_step_path(data_folder, first + p, first + p + 1),
                             np.array(res["R"][p]).reshape(3, 3).T, res["T"][p], res["state"][p], res["best_fit"][p],
                             (matches[p, :n].T + 1) if matches is not None else None)
    return res, matches, masks

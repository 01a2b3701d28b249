"""MATLAB-shaped mirror of the reference's function signatures for the hot path.

Same names, argument order, shapes (128 x K descriptors, 3 x N points), 1-based indices and
error behaviour as the reference .m / MEX files, so a test written against the reference reads
the same against this module.  Every function forwards to libpre3.so through
3pre_b200.api.Context -- nothing is computed in Python.  `M/` = /root/reference/matlab_code/.

New OPTIONAL trailing arguments (never required): explicit sample-index sets
(k x H, 1-based, as get_rand would have produced them), a seed, k, the adaptive-stop switch.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .api import Context, R2q, make_opts, make_frame_opts

_ctx = None


def context() -> Context:
    """Lazy one-time initialisation, like a MEX file's first call
    (M/mex_files/CorePar_Ver1/codegen/mex/corrcoef_partitioned/corrcoef_partitioned_mex.c:39-57)."""
    global _ctx
    if _ctx is None:
        _ctx = Context()
    return _ctx


def at_exit():
    """mexAtExit counterpart."""
    global _ctx
    if _ctx is not None:
        _ctx.close()
        _ctx = None


class MexError(RuntimeError):
    """mexErrMsgTxt: raised instead of longjmp-ing into the interpreter."""


def _cols(a, rows, dtype=None, name="argument"):
    """MATLAB rows x K matrix -> (K, rows) C-contiguous (same bytes as column-major)."""
    a = np.asarray(a) if dtype is None else np.asarray(a, dtype)
    if a.ndim != 2 or (rows is not None and a.shape[0] != rows):
        raise MexError(f"{name} must be a {rows} x K matrix")
    return np.ascontiguousarray(a.T)


def siftmatch(L1, L2, thresh=1.5, nargout=1):
    """matches = siftmatch(L1, L2[, thresh]);  [matches, D] = ...   (M/sift/siftmatch.m:1-12,
    gateway M/sift/siftmatch.c:139-250).  L1: ND x K1, L2: ND x K2, same class in
    {double, single, int8, uint8}.  matches: 2 x n double, 1-based, in k1 order."""
    if nargout > 2:
        raise MexError("Too many output arguments")  # siftmatch.c:156-158
    L1 = np.asarray(L1)
    L2 = np.asarray(L2)
    if (L1.ndim > 2 or L2.ndim > 2 or not np.issubdtype(L1.dtype, np.number)
            or not np.issubdtype(L2.dtype, np.number)):
        raise MexError("L1 and L2 must be two dimensional numeric arrays")  # :160-165
    L1 = np.atleast_2d(L1)
    L2 = np.atleast_2d(L2)
    if L1.shape[0] != L2.shape[0]:
        raise MexError("L1 and L2 must have the same number of rows")  # :171-173
    if L1.dtype != L2.dtype:
        raise MexError("L1 and L2 must be of the same class")  # :175-178
    if np.ndim(thresh) != 0 and np.size(thresh) != 1 or np.iscomplexobj(thresh):
        raise MexError("THRESH should be a real scalar")  # :183-186
    if L1.dtype not in (np.float64, np.float32, np.int8, np.uint8):
        raise MexError("Unsupported numeric class")  # :213-215
    thr = float(np.asarray(thresh).reshape(-1)[0])
    pairs, score = context().siftmatch(np.ascontiguousarray(L1.T), np.ascontiguousarray(L2.T), thr,
                                       want_score=nargout == 2)
    matches = (pairs.T + 1).astype(np.float64)  # 1-based [k1; k2]  (:241-242)
    if nargout == 2:
        return matches, score
    return matches


def siftmatch_sweep(descriptor1, descriptors2, thresh=1.5):
    """The loop of M/find_consistent_sift_matches.m:39-65: `matches = siftmatch(descriptor1, descriptor2)` for every
    later frame, the first frame's (good) descriptors fixed.  descriptor1: ND x K1; descriptors2: list of ND x K2_f
    matrices of the same class.  Returns the list of 2 x n double 1-based matches, one GPU call for all frames."""
    d1 = np.atleast_2d(np.asarray(descriptor1))
    ds = [np.atleast_2d(np.asarray(d)) for d in descriptors2]
    for d in ds:
        if d.shape[0] != d1.shape[0]:
            raise MexError("L1 and L2 must have the same number of rows")
        if d.dtype != d1.dtype:
            raise MexError("L1 and L2 must be of the same class")
    if d1.dtype not in (np.float64, np.float32, np.int8, np.uint8):
        raise MexError("Unsupported numeric class")
    if not ds:
        return []
    K2 = max(d.shape[1] for d in ds)
    L2 = np.zeros((len(ds), max(K2, 1), d1.shape[0]), d1.dtype)
    for f, d in enumerate(ds):
        L2[f, : d.shape[1]] = d.T
    out = context().siftmatch_sweep(np.ascontiguousarray(d1.T), L2, float(thresh),
                                    k2_count=np.array([d.shape[1] for d in ds], np.int32))
    return [(p.T + 1).astype(np.float64) for p, _ in out]


def matching_sift_based(SCAN_SIFT, features_info, step_global=0):
    """features_info = matching_sift_based(im, features_info, cam)  (M/matching_sift_based.m:27-151).  The reference
    loads SCAN_SIFT (fields Descriptor_RAW: ND x K2, SCALE_ORIENT_POS_RAW: >=2 x K2) from the step's .mat file; here
    it is an argument.  features_info: list of dicts (h, S, Descriptor, ...).  Matched and gated features get
    individually_compatible = 1, z, last_visible and the refreshed Descriptor (:125-129); everything else is left as
    it was.  Returns (features_info, discarded_sift_match)."""
    des2 = np.atleast_2d(np.asarray(_field(SCAN_SIFT, "Descriptor_RAW")))
    pos = np.asarray(_field(SCAN_SIFT, "SCALE_ORIENT_POS_RAW"), np.float64)
    idx = [i for i, fi in enumerate(features_info) if np.size(fi.get("h", ())) > 0]  # :108-113
    if not idx:
        return features_info, 0  # :114-116
    des1 = np.stack([np.asarray(features_info[i]["Descriptor"]).reshape(-1) for i in idx]).astype(des2.dtype)
    h = np.stack([np.asarray(features_info[i]["h"], np.float64).reshape(-1)[:2] for i in idx])
    S11 = np.array([np.asarray(features_info[i]["S"], np.float64)[0, 0] if np.size(features_info[i].get("S", ())) else
                    np.nan for i in idx])
    r = context().matching_sift_based_batch(des1[None], np.ascontiguousarray(des2.T)[None], h[None], S11[None],
                                            np.ascontiguousarray(pos[:2].T)[None])
    for j, i in enumerate(idx):
        if r["ic"][0, j]:
            k2 = int(r["match"][0, j])
            fi = features_info[i]
            fi["individually_compatible"] = 1
            fi["z"] = r["z"][0, j].copy()
            fi["last_visible"] = step_global
            fi["Descriptor"] = des2[:, k2].copy()
    return features_info, int(r["n_discarded"][0])


def find_transform_matrix(pset1, pset2):
    """[rot, trans, state] = find_transform_matrix(pset1, pset2)
    (M/mex_files/RANSAC_CALCULATION/find_transform_matrix.m:2-42); pset: 3 x n."""
    p1 = _cols(pset1, 3, np.float64, "pset1")
    p2 = _cols(pset2, 3, np.float64, "pset2")
    if p1.shape != p2.shape:
        raise MexError("pset1 and pset2 must have the same size")
    rot, tr, st = context().find_transform_matrix(p1, p2)
    return rot, tr.reshape(3, 1), st


def absoluteOrientationQuaternion(A, B, doScale=1):
    """[s, R, T, err] = absoluteOrientationQuaternion(A, B, doScale)
    (M/absoluteOrientationQuaternion.m:28-127); A, B: 3 x N, B ~ s*R*A + T.
    doScale defaults to 1 like nargin < 3 does in the reference (:32-34)."""
    A = np.asarray(A, np.float64)
    B = np.asarray(B, np.float64)
    if A.shape != B.shape:
        raise MexError("Point sets need to have same size.")  # :41-43
    if A.ndim != 2 or A.shape[0] != 3:
        raise MexError("Need points of dimension 3")  # :46-48
    if A.shape[1] < 4:
        raise MexError("Need at least 4 point pairs")  # :51-54
    s, R, T, err = context().horn(np.ascontiguousarray(A.T), np.ascontiguousarray(B.T), int(bool(doScale)))
    return s, R, T.reshape(3, 1), err


def RANSAC_CALC_VER2(Ya, Yb, options, Za=None, Zb=None, *, samples=None, seed=0, k=5, adaptive=True, H=None,
                     method="svd"):
    """[R, T, error, BestFit, State_RANSAC] = RANSAC_CALC_VER2(Ya, Yb, options[, Za, Zb])
    (M/mex_files/RANSAC_CALCULATION/RANSAC_CALC_VER2.m:2-201; method='horn' gives
    M/RANSAC_CALC_VER_test.m).  Ya, Yb: 3 x N.  options: dict / object with
    DistanceThreshold and MaxIteration (M/SIFT_match_save.m:49-50).  `error` is returned as
    the struct the reference ends up with (:199-201): {'ErrorSum', 'mYa', 'mYb'} (+ the
    pass-through 'mZa', 'mZb' when Za, Zb are given, :129-130).
    samples: k x H, 1-based (what get_rand(k, N) marks, ascending), optional."""
    ya = _cols(Ya, 3, np.float64, "Ya")
    yb = _cols(Yb, 3, np.float64, "Yb")
    if ya.shape != yb.shape:
        raise MexError("Ya and Yb must have the same size")
    get = (lambda n: options[n]) if isinstance(options, dict) else (lambda n: getattr(options, n))
    max_it = int(get("MaxIteration"))
    dthr = float(get("DistanceThreshold"))
    meth = L.METHOD_SVD if str(method).lower() == "svd" else L.METHOD_HORN
    s0 = None
    if samples is not None:
        s0 = np.ascontiguousarray(np.asarray(samples).T.astype(np.int32) - 1)  # (H,k) 0-based
        k = s0.shape[1]
    o = make_opts(method=meth, k=k, max_iteration=max_it, adaptive=adaptive,
                  H=(s0.shape[0] if s0 is not None else (H if H is not None else max_it)),
                  distance_threshold=dthr, seed=seed)
    r = context().ransac(ya, yb, s0, o)
    if r.status == 1:
        raise MexError("get_rand: not enough correspondences for a minimal sample")  # get_rand.m:39-41
    if r.status == 2:
        raise MexError("RANSAC_CALC_VER2: no hypothesis was recorded")
    err = {"ErrorSum": r.error_sum, "mYa": np.asarray(Ya, np.float64)[:, r.mask],
           "mYb": np.asarray(Yb, np.float64)[:, r.mask]}
    if Za is not None and Zb is not None:
        err["mZa"] = np.asarray(Za)[:, : len(r.mask)][:, r.mask]
        err["mZb"] = np.asarray(Zb)[:, : len(r.mask)][:, r.mask]
    return r.R, r.T.reshape(3, 1), err, r.best_fit, r.state


def SIFT_match_save(DataPre, DataCurrent, options=None, *, seed=0, pair_id=0, k=5, adaptive=True):
    """The compute of M/SIFT_match_save.m:19-53 for one pair (the .mat load/save stays with
    the caller).  DataPre / DataCurrent: SCAN_SIFT dicts with 'Descriptor' (128 x M) and
    'XYZ_DATA' (3 x M) (:8-14).  Returns the variables the reference saves (:79-80)."""
    options = options or {"DistanceThreshold": 0.05, "MaxIteration": 2000}  # :49-50
    d1 = _cols(DataPre["Descriptor"], None, None, "Descriptor")
    d2 = _cols(DataCurrent["Descriptor"], None, None, "Descriptor")
    x1 = _cols(DataPre["XYZ_DATA"], 3, np.float64, "XYZ_DATA")
    x2 = _cols(DataCurrent["XYZ_DATA"], 3, np.float64, "XYZ_DATA")
    o = make_opts(method=L.METHOD_SVD, k=k, max_iteration=int(options["MaxIteration"]), adaptive=adaptive,
                  H=int(options["MaxIteration"]), distance_threshold=float(options["DistanceThreshold"]), seed=seed)
    res, matches, masks = context().pairs(d1[None], d2[None], x1[None], x2[None], o, pair_id0=pair_id)
    from .api import unpack_result
    n = int(res["n_matches"][0])
    r = unpack_result(res[0], masks[0, :n].astype(bool))
    return {
        "matches": (matches[0, :n].T + 1).astype(np.float64),
        "R_RANSAC": r.R, "T_RANSAC": r.T.reshape(3, 1), "BestFit": r.best_fit, "State_RANSAC": r.state,
        "status": r.status, "PositionInliers": r.mask,
    }


def Calculate_V_Omega_RANSAC_my_version(DataPre, DataCurrent, **kw):
    """[T, q, R, State_RANSAC] as M/Calculate_V_Omega_RANSAC_my_version.m:9-34 returns them
    (q = R2q(R), [w x y z] with the slamToolbox sign convention)."""
    out = SIFT_match_save(DataPre, DataCurrent, **kw)
    return out["T_RANSAC"], R2q(out["R_RANSAC"]).reshape(4, 1), out["R_RANSAC"], out["State_RANSAC"]


# ---------------------------------------------------------------------------------------
# the step after the path: the EKF partial updates
# ---------------------------------------------------------------------------------------
def cov_est_RANSAC_deriv(Ya, Yb, R_ab, T_ab):
    """res = cov_est_RANSAC_deriv(Ya, Yb, R_ab, T_ab)  (M/cov_est_RANSAC_deriv.m:1-244; call site
    RANSAC_CALC_VER2.m:204-206, commented out in the reference).  Ya, Yb: 3 x n support set, Ya ~ R_ab*Yb + T_ab.
    Returns the reference's struct as a dict (sm_cov_censi 7 x 7 over [T; q]) plus the intermediate sums."""
    ya = _cols(Ya, 3, np.float64, "Ya")
    yb = _cols(Yb, 3, np.float64, "Yb")
    if ya.shape != yb.shape:
        raise MexError("Ya and Yb must have the same size")
    r = context().cov_est_ransac_batch(ya[None], yb[None], np.asarray(R_ab, np.float64).reshape(1, 3, 3),
                                       np.asarray(T_ab, np.float64).reshape(1, 3))[0]
    return {"sm_cov_censi": r["cov"], "Etot": r["Etot"], "Gtot": r["Gtot"], "G2tot": r["G2tot"], "dA_dz": r["dA_dz"],
            "s2": r["s2"]}


def update(x_km1_k, p_km1_k, H, R, z, h):
    """[x_k_k, p_k_k, K] = update(x_km1_k, p_km1_k, H, R, z, h)   (M/update.m:27-56).  H may be a scipy sparse
    matrix (the reference stacks sparse 2 x n blocks); K = 0 when z is empty (:54)."""
    if hasattr(H, "toarray"):
        H = H.toarray()
    z = np.asarray(z, np.float64).ravel()
    x = np.asarray(x_km1_k, np.float64)
    if z.size == 0:
        return x.copy(), np.array(p_km1_k, np.float64), 0
    xo, Po, K = context().ekf_update_dense(x.ravel(), p_km1_k, H, R, z, np.asarray(h, np.float64).ravel())
    return xo.reshape(x.shape), Po, K


# ---------------------------------------------------------------------------------------
# the step before the path: SR4000 frame -> filtered maps -> per-feature 3-D points
# ---------------------------------------------------------------------------------------
def _sr(sr_data):
    """rows x 176 MATLAB matrix -> (1,176,rows) C-contiguous (same bytes as column-major)."""
    a = np.asarray(sr_data, np.float64)
    if a.ndim != 2 or a.shape[1] < 176 or a.shape[0] not in (576, 720, 721):
        raise MexError("sr_data must be a 576/720/721 x 176 matrix (load of d1_%04d.dat)")
    return np.ascontiguousarray(a[:, :176].T)[None]


def read_xyz_sr4000(sr_data):
    """[x, y, z, confidence_map] = read_xyz_sr4000(prefix, k) from the matrix `load` returned (:3): z, x, y =
    imfilter(., fspecial('gaussian',[3 3],2), 'same') (M/read_xyz_sr4000.m:8-21); confidence_map raw (:26) or []."""
    a = np.asarray(sr_data, np.float64)
    x, y, z, _ = context().read_xyz_sr4000_batch(_sr(a), sigma=2.0, boundary=0)
    cm = a[576:720, :176].copy() if a.shape[0] >= 720 else np.zeros((0, 0))
    return x[0].T.copy(), y[0].T.copy(), z[0].T.copy(), cm


def read_sr4000_data_dr_ye(sr_data):
    """[x, y, z, confidence_map] of M/code_from_dr_ye/read_sr4000_data_dr_ye.m:8,27-36,88-90 (sigma 1, 'replicate');
    the amplitude-image output img1 (:11-25,42,70) feeds SIFT extraction and is not produced here."""
    a = np.asarray(sr_data, np.float64)
    x, y, z, _ = context().read_xyz_sr4000_batch(_sr(a), sigma=1.0, boundary=1, mode=1)
    cm = a[576:720, :176].copy() if a.shape[0] >= 720 else np.zeros((0, 0))
    return x[0].T.copy(), y[0].T.copy(), z[0].T.copy(), cm


def SIFT_extract_save(sr_data, frames, descriptors, idxScan=0):
    """The data part of M/SIFT_extract_save.m:44-88 for one scan, from the point where sift_vedal has returned
    (frames 4 x N with 0-based positions, descriptors 128 x N): SCAN_SIFT with Descriptor_RAW, SCALE_ORIENT_POS_RAW
    (positions + 1, :55-56), Descriptor, SCALE_ORIENT_POS, XYZ_DATA of the features that have valid 3-D data
    (inittialize_depth_my_version.m:40-85).  Image reading, SIFT extraction and the .mat save stay with the caller."""
    fr = _cols(frames, None, np.float64, "frames")
    de = _cols(descriptors, None, None, "descriptors")
    out = context().features_xyz_batch(_sr(sr_data), fr[None], desc=de[None], sigma=2.0, boundary=0, mode=0)
    if out["n_oob"]:
        raise MexError("Index exceeds matrix dimensions (inittialize_depth_my_version.m:40)")
    n = int(out["n_keep"][0])
    raw = np.array(frames, np.float64)
    raw[:2] += 1
    pos = out["frames_out"][0, :n].T.copy()
    pos[:2] += 1
    return {"idxScan": idxScan, "Descriptor_RAW": np.asarray(descriptors), "SCALE_ORIENT_POS_RAW": raw,
            "Descriptor": out["desc_out"][0, :n].T.copy(), "SCALE_ORIENT_POS": pos,
            "XYZ_DATA": out["xyz"][0, :n].T.copy(), "idxRemain": out["idx_remain"][0, :n] + 1}


def inittialize_depth_my_version(uvd, sr_data):
    """[initial_rho, Feature3d_in_code_coordinate] = inittialize_depth_my_version(uvd, step)
    (M/inittialize_depth_my_version.m:1-92) for ONE feature; uvd = [u; v] 1-based (column, row) as the callers pass
    it (SIFT_extract_save.m:77), the frame given as its sr_data matrix instead of a step number.  Empty outputs
    (None) when the reference returns []."""
    fr = np.array([[float(uvd[0]) - 1.0, float(uvd[1]) - 1.0]])
    out = context().features_xyz_batch(_sr(sr_data), fr[None], sigma=2.0, boundary=0, mode=0)
    if out["n_oob"]:
        raise MexError("Index exceeds matrix dimensions (inittialize_depth_my_version.m:40)")
    if not out["keep"][0, 0]:
        return None, None
    p = out["xyz_all"][0, 0]
    return 1.0 / np.sqrt(p @ p), p.copy()


# ---------------------------------------------------------------------------------------
# the code_from_dr_ye variant: what the live EKF calls (M/fv.m:47)
# ---------------------------------------------------------------------------------------
def _mround(v):
    """MATLAB round() for the positive pixel coordinates used here (half away from zero)."""
    return np.floor(np.asarray(v, np.float64) + 0.5).astype(np.int64)


def confidence_filtering(frm, des, confidence_map):
    """[frm, des] = confidence_filtering(frm, des, confidence_map)
    (M/code_from_dr_ye/confidence_filtering.m:1-13): drop features whose pixel confidence is below half the
    maximum.  Host-side index bookkeeping (K <= a few hundred lookups), no device work."""
    frm = np.asarray(frm, np.float64)
    cm = np.asarray(confidence_map)
    keep = cm[_mround(frm[1]) - 1, _mround(frm[0]) - 1] >= 0.5 * cm.max()
    return frm[:, keep], np.asarray(des)[:, keep]


def _pset(frm, idx, x, y, z):
    """pset(:,i) = [-x(ROW,COL); -y(ROW,COL); z(ROW,COL)] at the rounded frame position of feature idx(i)
    (M/code_from_dr_ye/ransac_dr_ye.m:13-19, vodometry_dr_ye.m:204-210)."""
    col = _mround(frm[0, idx]) - 1
    row = _mround(frm[1, idx]) - 1
    return np.stack([-np.asarray(x, np.float64)[row, col], -np.asarray(y, np.float64)[row, col],
                     np.asarray(z, np.float64)[row, col]])


def R2e(R):
    """e = [roll; pitch; yaw] (M/slamToolbox_11_02_18/FrameTransforms/Rotations/R2e.m:31-35)."""
    R = np.asarray(R, np.float64)
    return np.array([np.arctan2(R[2, 1], R[2, 2]), np.arcsin(-R[2, 0]), np.arctan2(R[1, 0], R[0, 0])])


def vodometry_dr_ye(Data1, Data2, *, confidence_map=False, samples=None, seed=0, pair_id=0, max_iteration=700):
    """[rot, phi, theta, psi, trans, error, pnum, op_num, sta, op_pset1, op_pset2, RANSAC_STAT] =
    vodometry_dr_ye(file1, file2) (M/code_from_dr_ye/vodometry_dr_ye.m:5-247) from the point where the two frames
    have been read and SIFT has run (:31,:67,:114,:116 stay with the caller: file I/O and SIFT extraction are not
    on this path).  Data: dict with 'frm' (4 x K as sift() returns it, 0-based pixel positions; :73-74 add 1),
    'des' (128 x K), 'x', 'y', 'z' (144 x 176 maps) and optionally 'confidence_map'.
    Descriptor matching (:139), the RANSAC iterations (:162-183 with ransac_dr_ye.m as the body), selection (:184),
    the refit (:211) and the residual statistics (:212-215) run in libpre3.so.
    samples: 4 x H, 1-based draws num_rs(1..4) per iteration (optional; else the seeded sampler)."""
    ctx = context()
    frm1 = np.array(Data1["frm"], np.float64)
    frm2 = np.array(Data2["frm"], np.float64)
    des1, des2 = np.asarray(Data1["des"]), np.asarray(Data2["des"])
    stat = {"nFeatures1": frm1.shape[1], "nFeatures2": frm2.shape[1], "nMatches": 0, "nIterationRansac": 0,
            "InlierRatio": 0, "nSupport": 0, "ErrorMean": 0, "ErrorStd": 0, "SolutionState": 0}
    frm1[:2] += 1  # :73-74
    frm2[:2] += 1  # :119-121
    if confidence_map:  # myCONFIG.FLAGS.CONFIDENCE_MAP (:80-82, :123-125)
        frm1, des1 = confidence_filtering(frm1, des1, Data1["confidence_map"])
        frm2, des2 = confidence_filtering(frm2, des2, Data2["confidence_map"])
    stat["nF1_Confidence_Filtered"], stat["nF2_Confidence_Filtered"] = frm1.shape[1], frm2.shape[1]
    match = siftmatch(des1, des2)  # :139
    pnum = match.shape[1]
    stat["nMatches"] = pnum
    fail = (np.zeros((3, 3)), 0.0, 0.0, 0.0, 0.0)
    if pnum < 4:  # :152-160
        stat["SolutionState"] = 4
        return (*fail, 1, pnum, 0, 0, np.zeros((3, 0)), np.zeros((3, 0)), stat)
    m0 = match.astype(np.int64) - 1
    pset1 = _pset(frm1, m0[0], Data1["x"], Data1["y"], Data1["z"])
    pset2 = _pset(frm2, m0[1], Data2["x"], Data2["y"], Data2["z"])
    s0 = None
    if samples is not None:
        s0 = np.ascontiguousarray(np.asarray(samples).T.astype(np.int32) - 1)[None]  # (1,H,4) 0-based
    res, masks, st, _ = ctx.vodometry_dr_ye_batch(np.ascontiguousarray(pset1.T)[None], np.ascontiguousarray(pset2.T)[None],
                                                  match=np.ascontiguousarray(m0.T.astype(np.int32))[None], samples=s0,
                                                  max_iteration=max_iteration,
                                                  H=(s0.shape[1] if s0 is not None else max_iteration), seed=seed)
    r, st = res[0], st[0]
    if r["status"] == 5:
        raise MexError("ransac_dr_ye: no point farther than 0.4 m (min of an empty set, ransac_dr_ye.m:21-22)")
    stat["nIterationRansac"] = int(st["n_iteration_ransac"])
    op_num = int(r["best_fit"])
    if r["status"] == 4:  # :187-194
        stat["SolutionState"] = 4
        return (*fail, 0, pnum, op_num, 0, np.zeros((3, 0)), np.zeros((3, 0)), stat)
    mask = masks[0, :pnum].astype(bool)
    op_pset1, op_pset2 = pset1[:, mask], pset2[:, mask]
    rot = np.array(r["R"]).reshape(3, 3).T.copy()
    trans = np.array(r["T"]).reshape(3, 1)
    sta = int(r["state"])
    stat.update(nSupport=op_num, ErrorMean=float(st["error_mean"]), ErrorStd=float(st["error_std"]), SolutionState=sta,
                GoodFrames1=frm1[:, m0[0, mask]], GoodDescriptor1=des1[:, m0[0, mask]],
                GoodFrames2=frm2[:, m0[1, mask]], GoodDescriptor2=des2[:, m0[1, mask]],
                InlierRatio=op_num / pnum * 100)
    if sta < 1:  # :230-235
        return rot, 0.0, 0.0, 0.0, 0.0, 2, pnum, op_num, sta, op_pset1, op_pset2, stat
    e = R2e(rot)  # :237-240
    return rot, e[0], e[1], e[2], trans, 0, pnum, op_num, sta, op_pset1, op_pset2, stat


def Calculate_V_Omega_RANSAC_dr_ye(Data1, Data2, **kw):
    """[T, q, R, sta, RANSAC_STAT] = Calculate_V_Omega_RANSAC_dr_ye(stepPre, stepCurrent)
    (M/code_from_dr_ye/Calculate_V_Omega_RANSAC_dr_ye.m:19-50; the result cache :12-32 stays with the caller)."""
    rot, _, _, _, trans, _, _, _, sta, _, _, stat = vodometry_dr_ye(Data1, Data2, **kw)
    if sta != 1:  # :41-44
        R = np.eye(3)
        return np.zeros((3, 1)), R2q(R).reshape(4, 1), R, sta, stat
    return trans, R2q(rot).reshape(4, 1), rot, sta, stat


# ---------------------------------------------------------------------------------------
# config 4: the 1-point-RANSAC EKF hypothesis path
# ---------------------------------------------------------------------------------------
StatData = {}  # the reference's global (M/ransac_hypotheses.m:34,84-85)


def compute_hypothesis_support_fast(xi, cam, state_vector_pattern, z_id, z_euc, threshold):
    """[hypothesis_support, positions_li_inliers_id, positions_li_inliers_euc] =
    compute_hypothesis_support_fast(xi, cam, state_vector_pattern, z_id, z_euc, threshold)
    (M/compute_hypothesis_support_fast.m:27-116).  xi: n x 1; pattern: n x 4; z_id: 2 x n_id or [];
    z_euc: 2 x n_euc or [].  Empty measurement sets give [] masks (:73-77,:112-116)."""
    xi = np.asarray(xi, np.float64).reshape(-1)
    z_id = np.asarray(z_id, np.float64)
    z_euc = np.asarray(z_euc, np.float64)
    zi = np.ascontiguousarray(z_id.T) if z_id.size else np.zeros((0, 2))
    ze = np.ascontiguousarray(z_euc.T) if z_euc.size else np.zeros((0, 2))
    try:
        sup, li, le = context().ekf_support(xi[None, :], cam, state_vector_pattern, zi, ze, float(threshold))
    except L.Pre3Error as e:
        if e.code == L.ERR_ARG:
            raise MexError(str(e)) from None
        raise
    return int(sup[0]), (li[0] if z_id.size else np.zeros(0, bool)), (le[0] if z_euc.size else np.zeros(0, bool))


def _field(obj, name):
    return obj[name] if isinstance(obj, dict) else getattr(obj, name)


def features_to_frame(filter, features_info, cam):
    """Unmarshal (filter, features_info, cam) the way ransac_hypotheses_mex.cpp does: per-feature
    type / state offset (generate_state_vector_pattern.m:30-51), measurement flags, z, h, the
    camera and feature blocks of H (calculate_Hi_inverse_depth_my_version.m:44-49 -- any other
    non-zero of H is an error) and R.  Returns the frames dict Context.ransac_hypotheses_batch takes."""
    x = np.asarray(_field(filter, "x_k_km1"), np.float64).reshape(-1)
    P = np.asarray(_field(filter, "p_k_km1"), np.float64)
    n, F = len(x), len(features_info)
    if P.shape != (n, n):
        raise MexError("p_k_km1 must be n x n")
    ty, pos = np.zeros(F, np.int32), np.zeros(F, np.int32)
    hz, ic, li = np.zeros(F, np.uint8), np.zeros(F, np.uint8), np.zeros(F, np.uint8)
    z, h = np.zeros((F, 2)), np.zeros((F, 2))
    Hc, Hf, R = np.zeros((F, 13, 2)), np.zeros((F, 6, 2)), np.zeros((F, 2, 2))
    p = 13
    for i, fi in enumerate(features_info):
        t = str(_field(fi, "type"))
        if t == "inversedepth":
            ty[i], nf = 0, 6
        elif t == "cartesian":
            ty[i], nf = 1, 3
        else:
            raise MexError("feature type must be 'inversedepth' or 'cartesian'")
        pos[i] = p
        p += nf
        zi = np.asarray(_field(fi, "z"), np.float64).reshape(-1)
        hz[i] = zi.size > 0
        ic[i] = bool(_field(fi, "individually_compatible"))
        li[i] = bool(fi.get("low_innovation_inlier", 0)) if isinstance(fi, dict) else bool(getattr(fi, "low_innovation_inlier", 0))
        if ic[i] or hz[i]:
            if zi.size:
                z[i] = zi[:2]
            h[i] = np.asarray(_field(fi, "h"), np.float64).reshape(-1)[:2]
            Hi = np.asarray(_field(fi, "H"), np.float64)
            if Hi.shape != (2, n):
                raise MexError("features_info(i).H must be 2 x n")
            rest = Hi.copy()
            rest[:, :13] = 0
            rest[:, pos[i]:pos[i] + nf] = 0
            if np.any(rest != 0):
                raise MexError("pre3:ekf: H has non-zeros outside the camera and the feature's own block")
            Hc[i] = Hi[:, :13].T
            Hf[i, :nf] = Hi[:, pos[i]:pos[i] + nf].T
            R[i] = np.asarray(_field(fi, "R"), np.float64).T
    if p != n:
        raise MexError("state size does not match the features (13 + 6 n_id + 3 n_euc)")
    return dict(x=x[None], P=np.ascontiguousarray(P.T)[None], type=ty[None], pos=pos[None], has_z=hz[None], ic=ic[None],
                li0=li[None], z=z[None], h=h[None], Hcam=Hc[None], Hfeat=Hf[None], R=R[None],
                std_z=float(_field(filter, "std_z")), cam=cam, n=n, F=F)


def predict_and_differentiate(x_k_k, cam, features_info):
    """features_info = predict_camera_measurements(x_k_k, cam, features_info);
    features_info = calculate_derivatives(x_k_k, cam, features_info)   (rescue_hi_inliers.m:32-33;
    M/predict_camera_measurements.m:27-68, M/calculate_derivatives.m:27-59).  cam needs f, Cx, Cy, k1, k2, nRows,
    nCols.  Updates h (1 x 2) and H (2 x n, dense) of every feature in place and returns the list."""
    x = np.asarray(x_k_k, np.float64).reshape(-1)
    F, n = len(features_info), len(x)
    ty, ps = np.zeros(F, np.int32), np.zeros(F, np.int32)
    p = 13
    for i, fi in enumerate(features_info):
        t = str(_field(fi, "type"))
        if t not in ("inversedepth", "cartesian"):
            raise MexError("feature type must be 'inversedepth' or 'cartesian'")
        ty[i], ps[i] = (0, p) if t == "inversedepth" else (1, p)
        p += 6 if ty[i] == 0 else 3
    if p != n:
        raise MexError("state size does not match the features (13 + 6 n_id + 3 n_euc)")
    has = np.array([np.size(fi.get("h", ())) > 0 for fi in features_info])
    h_in = np.zeros((F, 2))
    for i in np.flatnonzero(has):
        h_in[i] = np.asarray(features_info[i]["h"], np.float64).reshape(-1)[:2]
    h, has_o, _, Hc, Hf = context().ekf_predict_measurements_batch(x[None], cam, int(_field(cam, "nRows")),
                                                                   int(_field(cam, "nCols")), ty[None], ps[None],
                                                                   has[None], h_in[None])
    for i, fi in enumerate(features_info):
        if not has_o[0, i]:
            continue
        fi["h"] = h[0, i].reshape(1, 2).copy()
        Hi = np.zeros((2, n))
        Hi[:, :13] = Hc[0, i].T
        nf = 6 if ty[i] == 0 else 3
        Hi[:, ps[i]:ps[i] + nf] = Hf[0, i, :nf].T
        fi["H"] = Hi
    return features_info


def rescue_hi_inliers(filter, features_info, cam):
    """features_info = rescue_hi_inliers(filter, features_info, cam)   (M/@ekf_filter/rescue_hi_inliers.m:27-47).
    filter: dict / object with x_k_k, p_k_k.  Re-predicts h and H at x_k_k, then sets high_innovation_inlier for the
    features that are individually compatible but not low-innovation inliers (chi2inv(0.95, 2) = 5.9915)."""
    import torch
    x = np.asarray(_field(filter, "x_k_k"), np.float64).reshape(-1)
    P = np.asarray(_field(filter, "p_k_k"), np.float64)
    features_info = predict_and_differentiate(x, cam, features_info)
    F, n = len(features_info), len(x)
    ty, ps = np.zeros(F, np.int32), np.zeros(F, np.int32)
    ic, li = np.zeros(F, np.uint8), np.zeros(F, np.uint8)
    z, h, Hc, Hf = np.zeros((F, 2)), np.zeros((F, 2)), np.zeros((F, 13, 2)), np.zeros((F, 6, 2))
    p = 13
    for i, fi in enumerate(features_info):
        ty[i] = 0 if str(_field(fi, "type")) == "inversedepth" else 1
        ps[i] = p
        nf = 6 if ty[i] == 0 else 3
        p += nf
        ic[i] = bool(fi.get("individually_compatible", 0))
        li[i] = bool(fi.get("low_innovation_inlier", 0))
        if ic[i] and not li[i]:
            z[i] = np.asarray(fi["z"], np.float64).reshape(-1)[:2]
            h[i] = np.asarray(fi["h"], np.float64).reshape(-1)[:2]
            Hi = np.asarray(fi["H"], np.float64)
            Hc[i], Hf[i, :nf] = Hi[:, :13].T, Hi[:, ps[i]:ps[i] + nf].T
    ctx = context()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)[None]).cuda()
    frames = {"type": dev(ty), "pos": dev(ps), "ic": dev(ic), "z": dev(z), "h": dev(h), "Hcam": dev(Hc), "Hfeat": dev(Hf),
              "x": dev(x)}
    hi = torch.full((1, F), 7, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.ekf_rescue_hi_inliers_batch_dev(frames, dev(np.ascontiguousarray(P.T)), dev(li), hi)
    ctx.sync()
    out = hi[0].cpu().numpy()
    for i, fi in enumerate(features_info):
        if out[i] != 7:
            fi["high_innovation_inlier"] = int(out[i])
    return features_info


def ransac_hypotheses(filter, features_info, cam, *, selections=None, seed=0, n_hyp=1000, adaptive=True):
    """features_info = ransac_hypotheses(filter, features_info, cam)  (M/ransac_hypotheses.m:27-85).
    filter: dict / object with x_k_km1, p_k_km1, std_z (what get_x_k_km1 / get_p_k_km1 / get_std_z
    return); features_info: list of dicts (fields h, z, H, R, type, individually_compatible);
    cam: dict with f, Cx, Cy, k1, k2.  Returns a copy of features_info with low_innovation_inlier
    set like set_as_most_supported_hypothesis.m:32-53 and fills StatData (:84-85).
    selections (optional): 3 x H matrix of 1-based feature positions, one column per hypothesis,
    as select_random_match.m:58 would have returned them."""
    from .api import make_ekf_opts
    fr = features_to_frame(filter, features_info, cam)
    sel = None
    H = int(n_hyp)
    if selections is not None:
        s = np.asarray(selections)
        H = s.shape[1]
        sel = np.zeros((1, H, 3), np.int32)
        sel[0, :, :s.shape[0]] = s.T.astype(np.int32) - 1
    o = make_ekf_opts(n_hyp_init=n_hyp, H=H, adaptive=adaptive, seed=seed)
    res, li, _ = context().ransac_hypotheses_batch(fr, sel, o)
    if res["status"][0] == 1:
        raise MexError("select_random_match: Index exceeds matrix dimensions (no individually compatible match)")
    if res["status"][0] == 3:
        raise MexError("an individually compatible feature has no measurement z")
    out = []
    for i, fi in enumerate(features_info):
        g = dict(fi) if isinstance(fi, dict) else dict(vars(fi))
        if fr["has_z"][0, i] and res["best_hyp"][0] >= 0:
            g["low_innovation_inlier"] = int(li[0, i])
        out.append(g)
    StatData["RANSAC_ITER"] = float(res["n_hyp"][0])
    StatData["RANSAC_HYP_SUPPORT"] = int(res["max_support"][0])
    return out

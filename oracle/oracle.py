"""ctypes front-end of the CPU oracle (oracle/pre3_oracle*.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(3pre_b200/) never imports this module.

Array conventions mirror MATLAB: a "3 x N" matrix is passed as a numpy array of
shape (N, 3) C-contiguous (== 3 x N column-major); descriptors "128 x K" are
(K, 128) C-contiguous.  Rotations come back as 3x3 numpy arrays (row-major).
Indices are 0-based here; the MATLAB-facing layers add 1.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpre3_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle (and oracle/_ref when /root/reference exists)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.startswith("pre3_oracle") and f.endswith(".c")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)
    ref_so = os.path.join(_HERE, "_ref", "libsiftmatch_ref.so")
    if os.path.exists("/root/reference/matlab_code/sift/siftmatch.c") and (force or not os.path.exists(ref_so)):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


class RansacResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("state", C.c_int32),
        ("best_fit", C.c_int32),
        ("best_sample", C.c_int32),
        ("best_iter", C.c_int32),
        ("n_iter", C.c_int32),
        ("n_consumed", C.c_int32),
        ("pad", C.c_int32),
        ("thr", C.c_double),
        ("error_sum", C.c_double),
        ("R", C.c_double * 9),
        ("T", C.c_double * 3),
        ("R_hyp", C.c_double * 9),
        ("T_hyp", C.c_double * 3),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_distance_threshold.restype = C.c_double
        _lib.orc_adaptive_niter.restype = C.c_double
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


_MATCH = {
    np.dtype(np.float64): ("orc_siftmatch_f64", C.c_double),
    np.dtype(np.float32): ("orc_siftmatch_f32", C.c_float),
    np.dtype(np.int8): ("orc_siftmatch_i8", C.c_byte),
    np.dtype(np.uint8): ("orc_siftmatch_u8", C.c_ubyte),
}


def siftmatch(L1, L2, thresh: float = 1.5):
    """siftmatch.c:83-132.  L1:(K1,ND), L2:(K2,ND) same dtype.  Returns
    (pairs (n,2) int32 0-based, score (n,) float64)."""
    L1 = np.ascontiguousarray(L1)
    L2 = np.ascontiguousarray(L2)
    if L1.dtype != L2.dtype:
        raise ValueError("L1 and L2 must be of the same class")
    if L1.dtype not in _MATCH:
        raise ValueError("Unsupported numeric class")
    if L1.shape[1] != L2.shape[1]:
        raise ValueError("L1 and L2 must have the same number of rows")
    name, ct = _MATCH[L1.dtype]
    K1, ND = L1.shape
    K2 = L2.shape[0]
    pairs = np.zeros((max(K1, 1), 2), np.int32)
    score = np.zeros(max(K1, 1), np.float64)
    n = getattr(lib(), name)(_p(L1, ct), _p(L2, ct), K1, K2, ND, C.c_double(thresh), _p(pairs, C.c_int32), _p(score, C.c_double))
    return pairs[:n].copy(), score[:n].copy()


def matching_sift_based(des1, des2, h, S11, pos2, thresh: float = 1.5):
    """matching_sift_based.m:117-150 for one frame: siftmatch, then the search-region gate.  des1 (F,ND) descriptors of
    the predicted features, des2 (K2,ND), h (F,2), S11 (F,) with NaN for an empty S, pos2 (K2,2).  The reference reads
    S with the LOOP COUNTER over the matches (:120 `features_info(index_in_info(i)).S`), restated as is.  norm() of
    the 2-vector is sqrt(dx*dx + dy*dy).  Returns ic (F,) bool, z (F,2) NaN-filled, match (F,) 0-based | -1, n_match,
    n_discarded."""
    pairs, _ = siftmatch(des1, des2, thresh)
    F = des1.shape[0]
    ic, z, mt, disc = np.zeros(F, bool), np.full((F, 2), np.nan), np.full(F, -1, np.int32), 0
    for i, (k1, k2) in enumerate(pairs):
        S = S11[i]
        radius = 40.0 if np.isnan(S) else np.ceil(3.0 * np.sqrt(S))  # :121-127
        dx, dy = pos2[k2, 0] - h[k1, 0], pos2[k2, 1] - h[k1, 1]
        if np.sqrt(dx * dx + dy * dy) <= radius:  # :128-129
            ic[k1], z[k1], mt[k1] = True, pos2[k2], k2
        else:
            disc += 1  # :146
    return ic, z, mt, len(pairs), disc


def find_transform_matrix(pset1, pset2, idx=None):
    """find_transform_matrix.m:2-42.  pset: (n,3).  Returns rot(3,3), trans(3,), state."""
    p1, p2 = _f64(pset1), _f64(pset2)
    rot = np.zeros(9)
    tr = np.zeros(3)
    if idx is None:
        n, ip = p1.shape[0], None
    else:
        idx = np.ascontiguousarray(idx, np.int32)
        n, ip = idx.shape[0], _p(idx, C.c_int32)
    st = lib().orc_find_transform(_p(p1, C.c_double), _p(p2, C.c_double), ip, n, _p(rot, C.c_double), _p(tr, C.c_double))
    return rot.reshape(3, 3), tr, int(st)


def horn(A, B, doScale: int = 1, idx=None, allow_small: bool = False):
    """absoluteOrientationQuaternion.m:28-127.  A,B: (n,3); B ~ s*R*A + T.
    Returns s, R(3,3), T(3,), err.  Raises for n<4 like the reference (:51-54)."""
    a, b = _f64(A), _f64(B)
    if a.shape != b.shape:
        raise ValueError("Point sets need to have same size.")
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError("Need points of dimension 3")
    if idx is None:
        n, ip = a.shape[0], None
    else:
        idx = np.ascontiguousarray(idx, np.int32)
        n, ip = idx.shape[0], _p(idx, C.c_int32)
    s = C.c_double()
    err = C.c_double()
    R = np.zeros(9)
    T = np.zeros(3)
    rc = lib().orc_horn(_p(a, C.c_double), _p(b, C.c_double), ip, n, int(doScale), int(allow_small), C.byref(s), _p(R, C.c_double), _p(T, C.c_double), C.byref(err))
    if rc != 0:
        raise ValueError("Need at least 4 point pairs")
    return s.value, R.reshape(3, 3), T, err.value


def score(R, T, Ya, Yb, thr):
    """RANSAC_CALC_VER2.m:121-125,135.  Returns count, mask(N) bool, errsum."""
    R, T, ya, yb = _f64(R).reshape(9), _f64(T), _f64(Ya), _f64(Yb)
    N = ya.shape[0]
    mask = np.zeros(max(N, 1), np.uint8)
    es = C.c_double()
    c = lib().orc_score(_p(R, C.c_double), _p(T, C.c_double), _p(ya, C.c_double), _p(yb, C.c_double), N, C.c_double(thr), _p(mask, C.c_uint8), C.byref(es))
    return int(c), mask[:N].astype(bool), es.value


def distance_threshold(Yb):
    yb = _f64(Yb)
    return float(lib().orc_distance_threshold(_p(yb, C.c_double), yb.shape[0]))


def adaptive_niter(card, n_points, k=5, mult=5):
    return float(lib().orc_adaptive_niter(int(card), int(n_points), int(k), int(mult)))


def sample_sets(seed, pair, H, N, k):
    """(H,k) int32 0-based ascending subsets -- the seeded stand-in for get_rand.m."""
    out = np.zeros((H, k), np.int32)
    L = lib()
    L.orc_sample_set.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    for h in range(H):
        L.orc_sample_set(seed, pair, h, N, k, _p(out[h], C.c_int32))
    return out


@dataclass
class Ransac:
    status: int
    state: int
    best_fit: int
    best_sample: int
    best_iter: int
    n_iter: int
    n_consumed: int
    thr: float
    error_sum: float
    R: np.ndarray
    T: np.ndarray
    R_hyp: np.ndarray
    T_hyp: np.ndarray
    mask: np.ndarray
    counts: np.ndarray | None = None
    states: np.ndarray | None = None


def _unpack(res, mask, counts=None, states=None):
    return Ransac(
        res.status, res.state, res.best_fit, res.best_sample, res.best_iter, res.n_iter, res.n_consumed,
        res.thr, res.error_sum, np.array(res.R).reshape(3, 3), np.array(res.T), np.array(res.R_hyp).reshape(3, 3),
        np.array(res.T_hyp), mask, counts, states,
    )


def ransac(Ya, Yb, samples, method=0, max_iteration=2000, distance_threshold=0.05, adaptive=True):
    """RANSAC_CALC_VER2.m (method 0) / RANSAC_CALC_VER_test.m (method 1) with supplied
    sample sets (H,k) int32 0-based."""
    ya, yb = _f64(Ya), _f64(Yb)
    N = ya.shape[0]
    samples = np.ascontiguousarray(samples, np.int32)
    H, k = samples.shape
    res = RansacResult()
    mask = np.zeros(max(N, 1), np.uint8)
    counts = np.zeros(max(H, 1), np.int32)
    states = np.zeros(max(H, 1), np.int8)
    lib().orc_ransac(_p(ya, C.c_double), _p(yb, C.c_double), N, int(method), k, int(max_iteration), C.c_double(distance_threshold), int(bool(adaptive)), _p(samples, C.c_int32), H, C.byref(res), _p(mask, C.c_uint8), _p(counts, C.c_int32), _p(states, C.c_int8))
    return _unpack(res, mask[:N].astype(bool), counts[:H], states[:H])


def pair(desc1, desc2, xyz1, xyz2, seed, pair_id, H=2000, ratio=1.5, method=0, k=5, max_iteration=2000, distance_threshold=0.05, adaptive=True):
    """SIFT_match_save.m:33-53 for one pair (float64 descriptors).  Returns
    (matches (n,2) int32 0-based, Ransac)."""
    d1, d2, x1, x2 = _f64(desc1), _f64(desc2), _f64(xyz1), _f64(xyz2)
    K1, ND = d1.shape
    K2 = d2.shape[0]
    pairs = np.zeros((max(K1, 1), 2), np.int32)
    res = RansacResult()
    mask = np.zeros(max(K1, 1), np.uint8)
    L = lib()
    L.orc_pair.argtypes = [C.POINTER(C.c_double)] * 4 + [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(C.c_int32), C.POINTER(RansacResult), C.POINTER(C.c_uint8)]
    n = L.orc_pair(_p(d1, C.c_double), _p(d2, C.c_double), _p(x1, C.c_double), _p(x2, C.c_double), K1, K2, ND, ratio, method, k, max_iteration, distance_threshold, int(bool(adaptive)), seed, pair_id, H, _p(pairs, C.c_int32), C.byref(res), _p(mask, C.c_uint8))
    return pairs[:n].copy(), _unpack(res, mask[:n].astype(bool))


# ---------------------------------------------------------------------------------------------------
# code_from_dr_ye variant (oracle/pre3_oracle_dr_ye.c)
# ---------------------------------------------------------------------------------------------------
class DrYeResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("state", C.c_int32), ("op_num", C.c_int32), ("best_sample", C.c_int32),
                ("n_loops", C.c_int32), ("n_iteration_ransac", C.c_int32), ("pnum", C.c_int32), ("pad", C.c_int32),
                ("thr", C.c_double), ("error_sum", C.c_double), ("error_mean", C.c_double), ("error_std", C.c_double),
                ("R", C.c_double * 9), ("T", C.c_double * 3), ("R_hyp", C.c_double * 9), ("T_hyp", C.c_double * 3)]


@dataclass
class DrYe:
    status: int
    state: int
    op_num: int
    best_sample: int
    n_loops: int
    n_iteration_ransac: int
    pnum: int
    thr: float
    error_sum: float
    error_mean: float
    error_std: float
    R: np.ndarray
    T: np.ndarray
    R_hyp: np.ndarray
    T_hyp: np.ndarray
    mask: np.ndarray
    counts: np.ndarray


def dr_ye_sample(seed, pair, hyp, match, pnum):
    """ransac_dr_ye.m:28-48 on the seeded stream.  match (pnum,2) int32 or None.  Returns 4 0-based indices."""
    out = np.zeros(4, np.int32)
    mt = None if match is None else np.ascontiguousarray(match, np.int32)
    L = lib()
    L.orc_dr_ye_sample.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32), C.c_int,
                                   C.POINTER(C.c_int32)]
    L.orc_dr_ye_sample.restype = None
    L.orc_dr_ye_sample(int(seed), int(pair), int(hyp), _p(mt, C.c_int32) if mt is not None else None, int(pnum),
                       _p(out, C.c_int32))
    return out


def dr_ye_sample_stream(stream, match, n_sets):
    """The sampler on a recorded uniform stream; returns (sets (n_sets,4) 0-based, uniforms consumed)."""
    u = _f64(stream)
    mt = np.ascontiguousarray(match, np.int32)
    out = np.zeros((n_sets, 4), np.int32)
    L = lib()
    L.orc_dr_ye_sample_stream.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_int,
                                          C.POINTER(C.c_int32)]
    L.orc_dr_ye_sample_stream.restype = C.c_int
    used = L.orc_dr_ye_sample_stream(_p(u, C.c_double), u.size, _p(mt, C.c_int32), mt.shape[0], int(n_sets),
                                     _p(out, C.c_int32))
    return out, int(used)


def dr_ye_dist(Yb):
    yb = _f64(Yb)
    L = lib()
    L.orc_dr_ye_dist.argtypes = [C.POINTER(C.c_double), C.c_int]
    L.orc_dr_ye_dist.restype = C.c_double
    return float(L.orc_dr_ye_dist(_p(yb, C.c_double), yb.shape[0]))


def vodometry_dr_ye(Ya, Yb, match=None, samples=None, max_iteration=700, H=700, seed=0, pair=0):
    """RANSAC part of vodometry_dr_ye.m:147-220.  Ya = pset1, Yb = pset2 (N,3); match (N,2) | None;
    samples (H,4) 0-based draws | None (seeded)."""
    ya, yb = _f64(Ya), _f64(Yb)
    N = ya.shape[0]
    sm = None
    if samples is not None:
        sm = np.ascontiguousarray(samples, np.int32)
        H = sm.shape[0]
    mt = None if match is None else np.ascontiguousarray(match, np.int32)
    res = DrYeResult()
    mask = np.zeros(max(N, 1), np.uint8)
    counts = np.zeros(max(H, 1), np.int32)
    L = lib()
    L.orc_vodometry_dr_ye.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_int,
                                      C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_uint64, C.c_uint32,
                                      C.POINTER(DrYeResult), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)]
    L.orc_vodometry_dr_ye.restype = None
    L.orc_vodometry_dr_ye(_p(ya, C.c_double), _p(yb, C.c_double), _p(mt, C.c_int32) if mt is not None else None, N,
                          int(max_iteration), int(H), _p(sm, C.c_int32) if sm is not None else None, int(seed),
                          int(pair), C.byref(res), _p(mask, C.c_uint8), _p(counts, C.c_int32))
    return DrYe(res.status, res.state, res.op_num, res.best_sample, res.n_loops, res.n_iteration_ransac, res.pnum,
                res.thr, res.error_sum, res.error_mean, res.error_std, np.array(res.R).reshape(3, 3),
                np.array(res.T), np.array(res.R_hyp).reshape(3, 3), np.array(res.T_hyp), mask[:N].astype(bool),
                counts[:H].copy())


# ---------------------------------------------------------------------------------------------------
# frames -> per-feature 3-D points (oracle/pre3_oracle_frames.c)
# ---------------------------------------------------------------------------------------------------
def gaussian3(sigma):
    h = np.zeros(9)
    L = lib()
    L.orc_gaussian3.argtypes = [C.c_double, C.POINTER(C.c_double)]
    L.orc_gaussian3.restype = None
    L.orc_gaussian3(float(sigma), _p(h, C.c_double))
    return h.reshape(3, 3).T  # h[dr+1, dc+1]


def read_xyz(sr, sigma=2.0, boundary=0):
    """sr: (176, rows) C-contiguous = the rows x 176 column-major sr_data.  Returns x, y, z as (176,144) arrays
    (= 144 x 176 column-major)."""
    sr = _f64(sr)
    rows = sr.shape[1]
    x, y, z = (np.zeros((176, 144)) for _ in range(3))
    L = lib()
    L.orc_read_xyz.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_double, C.c_int] + [C.POINTER(C.c_double)] * 3
    L.orc_read_xyz.restype = None
    L.orc_read_xyz(_p(sr, C.c_double), rows, float(sigma), int(boundary), _p(x, C.c_double), _p(y, C.c_double),
                   _p(z, C.c_double))
    return x, y, z


def max_confidence(sr):
    sr = _f64(sr)
    L = lib()
    L.orc_max_confidence.argtypes = [C.POINTER(C.c_double), C.c_int]
    L.orc_max_confidence.restype = C.c_double
    return float(L.orc_max_confidence(_p(sr, C.c_double), sr.shape[1]))


def features_xyz(sr, frames, sigma=2.0, boundary=0, mode=0, use_conf=1):
    """sr: (176, rows); frames: (K, frame_ld).  Returns (xyz_all (K,3), keep (K,) bool, idx_remain (n,), n_oob)."""
    sr, fr = _f64(sr), _f64(frames)
    K, ld = fr.shape
    xyz = np.zeros((max(K, 1), 3))
    keep = np.zeros(max(K, 1), np.uint8)
    idx = np.zeros(max(K, 1), np.int32)
    oob = C.c_int32(0)
    L = lib()
    L.orc_features_xyz.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(C.c_double), C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint8),
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.orc_features_xyz.restype = C.c_int
    n = L.orc_features_xyz(_p(sr, C.c_double), sr.shape[1], float(sigma), int(boundary), int(mode), int(use_conf),
                           _p(fr, C.c_double), ld, K, _p(xyz, C.c_double), _p(keep, C.c_uint8), _p(idx, C.c_int32),
                           C.byref(oob))
    return xyz[:K], keep[:K].astype(bool), idx[:n].copy(), int(oob.value)


def cov_est_ransac_deriv(Ya, Yb, R, T):
    """cov_est_RANSAC_deriv.m (pre3_oracle_cov.c).  Ya, Yb (n,3), R (3,3), T (3,).  Returns dict like
    3pre_b200.api.unpack_cov."""
    ya, yb = _f64(Ya), _f64(Yb)
    r, t = _f64(np.asarray(R).reshape(3, 3)), _f64(np.asarray(T).reshape(3))
    cov, G2, G, dA, sc = np.zeros(49), np.zeros(49), np.zeros(7), np.zeros(42), np.zeros(2)
    f = lib().orc_cov_est_ransac_deriv
    f.restype = C.c_int
    st = f(_p(ya, C.c_double), _p(yb, C.c_double), C.c_int(ya.shape[0]), _p(r, C.c_double), _p(t, C.c_double),
           _p(cov, C.c_double), _p(G2, C.c_double), _p(G, C.c_double), _p(dA, C.c_double), _p(sc, C.c_double))
    return {"cov": cov.reshape(7, 7).T.copy(), "G2tot": G2.reshape(7, 7).T.copy(), "Gtot": G, "dA_dz": dA.reshape(6, 7).T.copy(),
            "Etot": float(sc[0]), "s2": float(sc[1]), "n": ya.shape[0], "status": 2 if st else 0}


def R2q(R):
    r = _f64(R).reshape(9)
    q = np.zeros(4)
    lib().orc_R2q(_p(r, C.c_double), _p(q, C.c_double))
    return q


# ---------------------------------------------------------------------------------------
# config 4: 1-point-RANSAC EKF hypotheses (oracle/pre3_oracle_ekf.c)
# ---------------------------------------------------------------------------------------
class EkfCam(C.Structure):
    _fields_ = [("f", C.c_double), ("Cx", C.c_double), ("Cy", C.c_double), ("k1", C.c_double), ("k2", C.c_double)]


class EkfResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_evaluated", C.c_int32), ("best_hyp", C.c_int32),
                ("max_support", C.c_int32), ("num_ic", C.c_int32), ("m", C.c_int32), ("n_hyp", C.c_double)]


def _cam(cam):
    return EkfCam(cam["f"], cam["Cx"], cam["Cy"], cam["k1"], cam["k2"])


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def sincos(x):
    """orc_sincos: the specified sin / cos pair (stand-in for MATLAB's in M/m.m:32-34)."""
    s, c = C.c_double(), C.c_double()
    lib().orc_sincos(C.c_double(float(x)), C.byref(s), C.byref(c))
    return s.value, c.value


def ekf_nhyp(support, num_ic):
    f = lib().orc_ekf_nhyp
    f.restype = C.c_double
    return f(int(support), int(num_ic))


def ekf_select(seed, frame, hyp, num_ic, m):
    out = np.zeros(3, np.int32)
    lib().orc_ekf_select(C.c_uint64(seed), C.c_uint32(frame), C.c_uint32(hyp), int(num_ic), int(m), _p(out, C.c_int32))
    return out[:m].copy()


def ekf_update(fr, sel):
    """xi = x + K (zi - hi) for the matches `sel` (M/ransac_hypotheses.m:51-63).  fr: synth_ekf.EkfFrame."""
    sel = _i32(sel)
    xi = np.zeros(fr.n)
    Pc = np.ascontiguousarray(fr.P.T)  # column-major
    lib().orc_ekf_update(_p(_f64(fr.x), C.c_double), _p(Pc, C.c_double), fr.n, _p(_i32(fr.pos), C.c_int32),
                         _p(_i32(fr.type), C.c_int32), _p(_f64(fr.z), C.c_double), _p(_f64(fr.h), C.c_double),
                         _p(_f64(fr.Hcam), C.c_double), _p(_f64(fr.Hfeat), C.c_double), _p(_f64(fr.R), C.c_double),
                         _p(sel, C.c_int32), len(sel), _p(xi, C.c_double))
    return xi


def pattern_to_lists(pattern):
    """logical pattern columns (generate_state_vector_pattern.m:29-51) -> 0-based index lists."""
    pattern = np.asarray(pattern)
    return [_i32(np.flatnonzero(pattern[:, c])) for c in range(4)]


def ekf_support(xi, cam, pattern, z_id, z_euc, threshold):
    """compute_hypothesis_support_fast.m:27-116.  z_id (n_id,2), z_euc (n_euc,2) [= 2 x n column-major].
    Returns support, li_id (bool), li_euc (bool), residuals."""
    ir, ia, irho, ixyz = pattern_to_lists(pattern)
    z_id, z_euc = _f64(z_id).reshape(-1, 2), _f64(z_euc).reshape(-1, 2)
    n_id, n_euc = len(z_id), len(z_euc)
    li_id, li_euc = np.zeros(n_id + 1, np.uint8), np.zeros(n_euc + 1, np.uint8)
    res = np.zeros(n_id + n_euc + 1)
    c = _cam(cam)
    sup = lib().orc_ekf_support(_p(_f64(xi), C.c_double), C.byref(c), _p(ir, C.c_int32), _p(ia, C.c_int32),
                                _p(irho, C.c_int32), _p(z_id, C.c_double), n_id, _p(ixyz, C.c_int32),
                                _p(z_euc, C.c_double), n_euc, C.c_double(threshold), _p(li_id, C.c_uint8),
                                _p(li_euc, C.c_uint8), _p(res, C.c_double))
    return sup, li_id[:n_id].astype(bool), li_euc[:n_euc].astype(bool), res[:n_id + n_euc].copy()


def ekf_predict(x, cam, n_rows, n_cols, types, pos, has_h, h_in):
    """predict_camera_measurements + calculate_derivatives at x (rescue_hi_inliers.m:32-33; pre3_oracle_ekf.c:
    orc_ekf_predict).  Returns h (F,2), has_h (F,) bool, predicted (F,) bool, Hcam (F,13,2), Hfeat (F,6,2)."""
    x = _f64(x)
    ty, ps = _i32(types), _i32(pos)
    F = len(ty)
    hh, hi = np.ascontiguousarray(np.asarray(has_h).astype(np.uint8)), _f64(h_in).reshape(F, 2)
    h, has, pred = np.zeros((F, 2)), np.zeros(F, np.uint8), np.zeros(F, np.uint8)
    Hc, Hf = np.zeros((F, 13, 2)), np.zeros((F, 6, 2))
    c = _cam(cam)
    lib().orc_ekf_predict(_p(x, C.c_double), len(x), C.byref(c), int(n_rows), int(n_cols), F, _p(ty, C.c_int32),
                          _p(ps, C.c_int32), _p(hh, C.c_uint8), _p(hi, C.c_double), _p(h, C.c_double),
                          _p(has, C.c_uint8), _p(pred, C.c_uint8), _p(Hc, C.c_double), _p(Hf, C.c_double))
    return h, has.astype(bool), pred.astype(bool), Hc, Hf


def ransac_hypotheses(fr, sel=None, H=1000, n_hyp_init=1000, seed=0, frame_id=0, adaptive=True):
    """ransac_hypotheses.m:27-85 on one synth_ekf.EkfFrame.  sel: (H,3) 0-based feature indices or None
    (seeded).  Returns dict(li, n_hyp, max_support, best_hyp, n_evaluated, supports, status, m, num_ic)."""
    if sel is not None:
        sel = _i32(sel)
        H = len(sel)
    li = np.ascontiguousarray(fr.li0, dtype=np.uint8).copy()
    out = EkfResult()
    supports = np.full(max(H, 1), -1, np.int32)
    c = _cam(fr.cam)
    Pc = np.ascontiguousarray(fr.P.T)
    hz, ic = np.ascontiguousarray(fr.has_z, np.uint8), np.ascontiguousarray(fr.ic, np.uint8)
    lib().orc_ransac_hypotheses(_p(_f64(fr.x), C.c_double), _p(Pc, C.c_double), fr.n, C.c_double(fr.std_z), C.byref(c),
                                fr.F, _p(_i32(fr.type), C.c_int32), _p(_i32(fr.pos), C.c_int32), _p(hz, C.c_uint8),
                                _p(ic, C.c_uint8), _p(_f64(fr.z), C.c_double), _p(_f64(fr.h), C.c_double),
                                _p(_f64(fr.Hcam), C.c_double), _p(_f64(fr.Hfeat), C.c_double), _p(_f64(fr.R), C.c_double),
                                _p(sel, C.c_int32) if sel is not None else None, int(H), int(n_hyp_init),
                                int(bool(adaptive)), C.c_uint64(seed), C.c_uint32(frame_id), _p(li, C.c_uint8), C.byref(out),
                                _p(supports, C.c_int32))
    return dict(li=li, n_hyp=out.n_hyp, max_support=out.max_support, best_hyp=out.best_hyp,
                n_evaluated=out.n_evaluated, supports=supports[:out.n_evaluated].copy(), status=out.status, m=out.m,
                num_ic=out.num_ic)

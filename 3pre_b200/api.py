"""Host side of libpre3.so: a thin object over the C ABI (include/pre3.h).

Array conventions (shared with oracle/oracle.py so that parity tests compare like with like):
a MATLAB "3 x N" matrix is a numpy array of shape (N, 3), C-contiguous (the same bytes as
3 x N column-major); "128 x K" descriptors are (K, 128); rotations are returned as 3x3
numpy arrays in the usual row/column sense; indices are 0-based.  The MATLAB-shaped,
1-based mirror of the reference's function signatures is 3pre_b200/matlab.py.

torch is used only for device memory / streams (the *_dev methods take CUDA tensors).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from ._lib import Pre3Error, RansacOpts, PairResult, Cam, EkfOpts, EkfResult, DrYeStat, FrameOpts

_CLS = {
    np.dtype(np.float64): L.CLASS_DOUBLE,
    np.dtype(np.float32): L.CLASS_SINGLE,
    np.dtype(np.int8): L.CLASS_INT8,
    np.dtype(np.uint8): L.CLASS_UINT8,
}

RESULT_DTYPE = np.dtype(
    [("status", "<i4"), ("state", "<i4"), ("best_fit", "<i4"), ("best_sample", "<i4"), ("best_iter", "<i4"),
     ("n_iter", "<i4"), ("n_consumed", "<i4"), ("n_matches", "<i4"), ("thr", "<f8"), ("error_sum", "<f8"),
     ("R", "<f8", (9,)), ("T", "<f8", (3,)), ("R_hyp", "<f8", (9,)), ("T_hyp", "<f8", (3,))]
)
assert RESULT_DTYPE.itemsize == C.sizeof(PairResult) == 240
EKF_RESULT_DTYPE = np.dtype([("status", "<i4"), ("n_evaluated", "<i4"), ("best_hyp", "<i4"), ("max_support", "<i4"),
                             ("num_ic", "<i4"), ("m", "<i4"), ("n_hyp", "<f8")])
assert EKF_RESULT_DTYPE.itemsize == C.sizeof(EkfResult) == 32
DR_YE_STAT_DTYPE = np.dtype([("error_mean", "<f8"), ("error_std", "<f8"), ("dist", "<f8"),
                             ("n_iteration_ransac", "<i4"), ("n_loops", "<i4")])
assert DR_YE_STAT_DTYPE.itemsize == C.sizeof(DrYeStat) == 32
COV_RESULT_DTYPE = np.dtype([("cov", "<f8", (49,)), ("G2tot", "<f8", (49,)), ("Gtot", "<f8", (7,)), ("dA_dz", "<f8", (42,)),
                             ("Etot", "<f8"), ("s2", "<f8"), ("n", "<i4"), ("status", "<i4")])
assert COV_RESULT_DTYPE.itemsize == 8 * (49 + 49 + 7 + 42 + 2) + 8


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch tensor


def _c(a, dtype=None):
    return np.ascontiguousarray(a, dtype=dtype)


def unpack_cov(rec) -> dict:
    """One COV_RESULT_DTYPE record -> MATLAB-shaped arrays."""
    return {"cov": rec["cov"].reshape(7, 7).T.copy(), "G2tot": rec["G2tot"].reshape(7, 7).T.copy(),
            "Gtot": rec["Gtot"].copy(), "dA_dz": rec["dA_dz"].reshape(6, 7).T.copy(), "Etot": float(rec["Etot"]),
            "s2": float(rec["s2"]), "n": int(rec["n"]), "status": int(rec["status"])}


def make_frame_opts(sigma=2.0, boundary=0, mode=0, rows=720, use_confidence=1) -> FrameOpts:
    """Defaults = M/read_xyz_sr4000.m:8-21 + M/inittialize_depth_my_version.m; the code_from_dr_ye flavour is
    sigma=1, boundary=1 ('replicate'), mode=1."""
    return FrameOpts(float(sigma), int(boundary), int(mode), int(rows), int(bool(use_confidence)))


def make_opts(method=L.METHOD_SVD, k=5, max_iteration=2000, adaptive=True, H=2000, distance_threshold=0.05,
              ratio=1.5, seed=0) -> RansacOpts:
    """options struct of RANSAC_CALC_VER2 (DistanceThreshold, MaxIteration: SIFT_match_save.m:49-50)
    plus the explicit knobs the reference hard-codes (k: RANSAC_CALC_VER2.m:85)."""
    return RansacOpts(int(method), int(k), int(max_iteration), int(bool(adaptive)), int(H), 0,
                      float(distance_threshold), float(ratio), int(seed))


def make_cam(cam) -> Cam:
    """cam struct fields used on the path (M/initialize_cam.m:64-76); accepts a dict or a Cam."""
    if isinstance(cam, Cam):
        return cam
    return Cam(float(cam["f"]), float(cam["Cx"]), float(cam["Cy"]), float(cam["k1"]), float(cam["k2"]))


def make_ekf_opts(n_hyp_init=1000, H=1000, adaptive=True, seed=0) -> EkfOpts:
    """n_hyp_init: ransac_hypotheses.m:35; H: selections available; adaptive: the stop rule :77-80."""
    return EkfOpts(int(n_hyp_init), int(H), int(bool(adaptive)), 0, int(seed))


@dataclass
class Ransac:
    """One pair's outputs, same field meaning as oracle.oracle.Ransac."""
    status: int
    state: int
    best_fit: int
    best_sample: int
    best_iter: int
    n_iter: int
    n_consumed: int
    n_matches: int
    thr: float
    error_sum: float
    R: np.ndarray
    T: np.ndarray
    R_hyp: np.ndarray
    T_hyp: np.ndarray
    mask: np.ndarray | None = None
    counts: np.ndarray | None = None
    states: np.ndarray | None = None


def unpack_result(rec, mask=None, counts=None, states=None) -> Ransac:
    """rec: one element of a RESULT_DTYPE array.  R is stored column-major in the ABI."""
    return Ransac(int(rec["status"]), int(rec["state"]), int(rec["best_fit"]), int(rec["best_sample"]),
                  int(rec["best_iter"]), int(rec["n_iter"]), int(rec["n_consumed"]), int(rec["n_matches"]),
                  float(rec["thr"]), float(rec["error_sum"]), np.array(rec["R"]).reshape(3, 3).T.copy(),
                  np.array(rec["T"]), np.array(rec["R_hyp"]).reshape(3, 3).T.copy(), np.array(rec["T_hyp"]),
                  mask, counts, states)


class Context:
    """pre3_ctx: one CUDA device, one stream, one workspace arena."""

    def __init__(self, device: int = -1):
        self._lib = L.load()
        h = C.c_void_p()
        rc = self._lib.pre3_create(C.byref(h), int(device))
        self._h = h
        if rc != L.OK:
            msg = self._lib.pre3_last_error(h).decode() if h else "pre3_create failed"
            if h:
                self._lib.pre3_destroy(h)
            self._h = None
            raise Pre3Error(rc, msg)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pre3_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != L.OK:
            raise Pre3Error(rc, self._lib.pre3_last_error(self._h).decode())

    # ---- context knobs ------------------------------------------------------------------
    @staticmethod
    def version() -> str:
        return L.load().pre3_version().decode()

    def set_stream(self, cuda_stream_ptr: int):
        self._ck(self._lib.pre3_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def use_torch_stream(self):
        import torch
        self.set_stream(torch.cuda.current_stream().cuda_stream)

    def set_match_engine(self, engine: int):
        self._ck(self._lib.pre3_set_match_engine(self._h, int(engine)))

    def set_graphs(self, on: bool = True):
        """CUDA-graph replay of pairs_dev / sequence_dev calls that repeat a signature (pre3_set_graphs)."""
        self._ck(self._lib.pre3_set_graphs(self._h, int(bool(on))))

    def set_pipeline(self, chunks: int = -1):
        """Chunked stage pipeline of pairs_dev / sequence_dev (pre3_set_pipeline): -1 automatic, 0 off, n chunks."""
        self._ck(self._lib.pre3_set_pipeline(self._h, int(chunks)))

    def sync(self):
        self._ck(self._lib.pre3_sync(self._h))

    def launch_count(self) -> int:
        return int(self._lib.pre3_launch_count(self._h))

    def transfer_bytes(self):
        """(host->device, device->host) bytes moved by pairs() so far."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._ck(self._lib.pre3_transfer_bytes(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def eval_schedule(self, opts: RansacOpts, P: int | None = None):
        """Wave boundaries of the hypothesis evaluation (pre3_eval_schedule; with P: for a call of exactly P pairs)."""
        ends = np.zeros(40, np.int32)
        if P is None:
            n = self._lib.pre3_eval_schedule(C.byref(opts), _ptr(ends), 40)
        else:
            n = self._lib.pre3_eval_schedule_for(C.byref(opts), int(P), _ptr(ends), 40)
        if n < 0:
            raise L.Pre3Error(n, "pre3_eval_schedule")
        return ends[:n].copy()

    def timing_enable(self, on: bool = True):
        self._ck(self._lib.pre3_timing_enable(self._h, int(bool(on))))

    def timing_read(self) -> dict:
        """{category: (total ms, launches)} of the bracketed launches since the last read."""
        ms = np.zeros(L.TIMING_NCAT)
        cnt = np.zeros(L.TIMING_NCAT, np.int64)
        self._ck(self._lib.pre3_timing_read(self._h, _ptr(ms), _ptr(cnt)))
        return {self._lib.pre3_timing_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(L.TIMING_NCAT)
                if cnt[i]}

    def measure_tmem_read(self) -> float:
        """GB/s of tcgen05.ld over the whole chip."""
        v = C.c_double(0.0)
        self._ck(self._lib.pre3_measure_tmem_read(self._h, C.byref(v)))
        return v.value

    def measure_fp64_peak(self) -> float:
        v = C.c_double(0.0)
        self._ck(self._lib.pre3_measure_fp64_peak(self._h, C.byref(v)))
        return v.value

    def ekf_eval_schedule(self, opts: EkfOpts):
        ends = np.zeros(40, np.int32)
        n = self._lib.pre3_ekf_eval_schedule(C.byref(opts), _ptr(ends), 40)
        if n < 0:
            raise Pre3Error(n, "pre3_ekf_eval_schedule")
        return ends[:n].copy()

    def measure_fp32_peak_3reg(self) -> float:
        v = C.c_double(0.0)
        self._ck(self._lib.pre3_measure_fp32_peak_mode(self._h, 1, C.byref(v)))
        return v.value

    def measure_fp32_peak(self) -> float:
        """FFMA-chain microbenchmark, TFLOP/s (the scoring roofline's denominator)."""
        t = np.zeros(1)
        self._ck(self._lib.pre3_measure_fp32_peak(self._h, _ptr(t)))
        return float(t[0])

    # ---- stage 1 ------------------------------------------------------------------------
    def siftmatch(self, L1, L2, thresh: float = 1.5, want_score: bool = True):
        """L1 (K1,ND), L2 (K2,ND) same dtype in {f64,f32,i8,u8}.  Returns (pairs (n,2) int32
        0-based in k1 order, score (n,) float64 | None) -- siftmatch.c:83-132.  want_score=False
        is `matches = siftmatch(L1,L2)` (nout == 1): the library may then skip the exact
        distance of rows whose acceptance is already certified."""
        L1, L2 = _c(L1), _c(L2)
        if L1.dtype != L2.dtype:
            raise ValueError("L1 and L2 must be of the same class")
        if L1.dtype not in _CLS:
            raise ValueError("Unsupported numeric class")
        if L1.ndim != 2 or L2.ndim != 2 or L1.shape[1] != L2.shape[1]:
            raise ValueError("L1 and L2 must have the same number of rows")
        K1, ND = L1.shape
        K2 = L2.shape[0]
        pairs = np.zeros((max(K1, 1), 2), np.int32)
        score = np.zeros(max(K1, 1), np.float64) if want_score else None
        n = np.zeros(1, np.int32)
        self._ck(self._lib.pre3_siftmatch(self._h, _ptr(L1), _ptr(L2), _CLS[L1.dtype], K1, K2, ND, float(thresh),
                                          _ptr(pairs), _ptr(score), _ptr(n)))
        return pairs[: n[0]].copy(), (score[: n[0]].copy() if want_score else None)

    def siftmatch_batch(self, L1, L2, thresh: float = 1.5, k1_count=None, k2_count=None):
        """L1 (P,K1,ND), L2 (P,K2,ND).  Returns list of (pairs, score) per problem."""
        L1, L2 = _c(L1), _c(L2)
        if L1.dtype != L2.dtype or L1.dtype not in _CLS:
            raise ValueError("Unsupported numeric class")
        P, K1, ND = L1.shape
        K2 = L2.shape[1]
        pairs = np.zeros((P, max(K1, 1), 2), np.int32)
        score = np.zeros((P, max(K1, 1)), np.float64)
        n = np.zeros(max(P, 1), np.int32)
        k1c = None if k1_count is None else _c(k1_count, np.int32)
        k2c = None if k2_count is None else _c(k2_count, np.int32)
        self._ck(self._lib.pre3_siftmatch_batch(self._h, _ptr(L1), _ptr(L2), _CLS[L1.dtype], P, K1, K2, ND,
                                                _ptr(k1c), _ptr(k2c), float(thresh), _ptr(pairs), _ptr(score),
                                                _ptr(n)))
        return [(pairs[p, : n[p]].copy(), score[p, : n[p]].copy()) for p in range(P)]

    def siftmatch_sweep(self, L1, L2, thresh: float = 1.5, k2_count=None):
        """One set L1 (K1,ND) against P sets L2 (P,K2,ND) -- the loop of find_consistent_sift_matches.m:39-65.
        Returns list of (pairs, score) per later frame."""
        L1, L2 = _c(L1), _c(L2)
        if L1.dtype != L2.dtype or L1.dtype not in _CLS:
            raise ValueError("Unsupported numeric class")
        K1, ND = L1.shape
        P, K2 = L2.shape[0], L2.shape[1]
        pairs = np.zeros((max(P, 1), max(K1, 1), 2), np.int32)
        score = np.zeros((max(P, 1), max(K1, 1)), np.float64)
        n = np.zeros(max(P, 1), np.int32)
        k2c = None if k2_count is None else _c(k2_count, np.int32)
        self._ck(self._lib.pre3_siftmatch_sweep(self._h, _ptr(L1), _ptr(L2), _CLS[L1.dtype], P, K1, K2, ND, _ptr(k2c),
                                                float(thresh), _ptr(pairs), _ptr(score), _ptr(n)))
        return [(pairs[p, : n[p]].copy(), score[p, : n[p]].copy()) for p in range(P)]

    def siftmatch_sweep_dev(self, L1, L2, pairs, score, n_out, thresh=1.5, k2_count=None):
        import torch
        cls = {torch.float64: L.CLASS_DOUBLE, torch.float32: L.CLASS_SINGLE, torch.int8: L.CLASS_INT8,
               torch.uint8: L.CLASS_UINT8}[L1.dtype]
        K1, ND = L1.shape
        P, K2 = L2.shape[0], L2.shape[1]
        self._ck(self._lib.pre3_siftmatch_sweep_dev(self._h, _ptr(L1), _ptr(L2), cls, P, K1, K2, ND, _ptr(k2_count),
                                                    float(thresh), _ptr(pairs), _ptr(score), _ptr(n_out)))

    def matching_sift_based_batch(self, des1, des2, h, S11, pos2, f_count=None, k2_count=None, thresh: float = 1.5):
        """matching_sift_based.m:104-135 for P frames.  des1 (P,F,ND) descriptors of the predicted features, des2
        (P,K2,ND) Descriptor_RAW, h (P,F,2), S11 (P,F) (NaN = empty S), pos2 (P,K2,2).  Returns a dict: ic (P,F) bool,
        z (P,F,2) (NaN where not compatible), match (P,F) 0-based k2 | -1, n_match (P,), n_discarded (P,)."""
        d1, d2 = _c(des1), _c(des2)
        if d1.dtype != d2.dtype or d1.dtype not in _CLS:
            raise ValueError("Unsupported numeric class")
        P, F, ND = d1.shape
        K2 = d2.shape[1]
        h, S11, pos2 = _c(h, np.float64), _c(S11, np.float64), _c(pos2, np.float64)
        Pm, Fm = max(P, 1), max(F, 1)
        ic, z, mt = np.zeros((Pm, Fm), np.uint8), np.zeros((Pm, Fm, 2)), np.zeros((Pm, Fm), np.int32)
        nm, nd = np.zeros(Pm, np.int32), np.zeros(Pm, np.int32)
        fc = None if f_count is None else _c(f_count, np.int32)
        k2c = None if k2_count is None else _c(k2_count, np.int32)
        self._ck(self._lib.pre3_matching_sift_based_batch(self._h, _ptr(d1), _ptr(d2), _CLS[d1.dtype], P, F, K2, ND,
                                                          _ptr(fc), _ptr(k2c), _ptr(h), _ptr(S11), _ptr(pos2),
                                                          float(thresh), _ptr(ic), _ptr(z), _ptr(mt), _ptr(nm),
                                                          _ptr(nd)))
        return {"ic": ic[:P, :F].astype(bool), "z": z[:P, :F], "match": mt[:P, :F], "n_match": nm[:P],
                "n_discarded": nd[:P]}

    # ---- stage 2 ------------------------------------------------------------------------
    def find_transform_matrix(self, pset1, pset2):
        """pset (n,3).  Returns rot (3,3), trans (3,), state (find_transform_matrix.m:2-42)."""
        p1, p2 = _c(pset1, np.float64), _c(pset2, np.float64)
        rot, tr, st = np.zeros(9), np.zeros(3), np.zeros(1, np.int32)
        self._ck(self._lib.pre3_find_transform_matrix(self._h, _ptr(p1), _ptr(p2), p1.shape[0], _ptr(rot), _ptr(tr),
                                                      _ptr(st)))
        return rot.reshape(3, 3).T.copy(), tr, int(st[0])

    def horn(self, A, B, do_scale: int = 1):
        """A,B (n,3), B ~ s*R*A + T.  Returns s, R, T, err (absoluteOrientationQuaternion.m:28-127)."""
        a, b = _c(A, np.float64), _c(B, np.float64)
        if a.shape != b.shape:
            raise ValueError("Point sets need to have same size.")
        if a.ndim != 2 or a.shape[1] != 3:
            raise ValueError("Need points of dimension 3")
        if a.shape[0] < 4:
            raise ValueError("Need at least 4 point pairs")
        s, err, R, T = np.zeros(1), np.zeros(1), np.zeros(9), np.zeros(3)
        self._ck(self._lib.pre3_horn(self._h, _ptr(a), _ptr(b), a.shape[0], int(do_scale), _ptr(s), _ptr(R), _ptr(T),
                                     _ptr(err)))
        return float(s[0]), R.reshape(3, 3).T.copy(), T, float(err[0])

    def fit_batch(self, Ya, Yb, samples, method=L.METHOD_SVD):
        """samples (H,k) int32 0-based.  Returns R (H,3,3), T (H,3), state (H,)."""
        ya, yb = _c(Ya, np.float64), _c(Yb, np.float64)
        s = _c(samples, np.int32)
        H, k = s.shape
        R, T, st = np.zeros((H, 9)), np.zeros((H, 3)), np.zeros(H, np.int32)
        self._ck(self._lib.pre3_fit_batch(self._h, _ptr(ya), _ptr(yb), ya.shape[0], _ptr(s), k, H, int(method),
                                          _ptr(R), _ptr(T), _ptr(st)))
        return R.reshape(H, 3, 3).transpose(0, 2, 1).copy(), T, st

    # ---- stage 3 ------------------------------------------------------------------------
    def score_batch(self, R, T, Ya, Yb, thr, want_errsum=True, want_mask=True):
        """R (H,3,3), T (H,3).  Returns count (H,), errsum (H,)|None, mask (H,N) bool|None."""
        R = np.asarray(R, np.float64).reshape(-1, 3, 3)
        H = R.shape[0]
        Rc = _c(R.transpose(0, 2, 1)).reshape(H, 9)  # column-major per hypothesis
        T = _c(np.asarray(T, np.float64).reshape(H, 3))
        ya, yb = _c(Ya, np.float64), _c(Yb, np.float64)
        N = ya.shape[0]
        cnt = np.zeros(H, np.int32)
        es = np.zeros(H) if want_errsum else None
        mk = np.zeros((H, max(N, 1)), np.uint8) if want_mask else None
        self._ck(self._lib.pre3_score_batch(self._h, _ptr(Rc), _ptr(T), H, _ptr(ya), _ptr(yb), N, float(thr),
                                            _ptr(cnt), _ptr(es), _ptr(mk)))
        if mk is not None:
            mk = mk.reshape(-1)[: H * N].reshape(H, N).astype(bool)
        return cnt, es, mk

    # ---- stages 2-4 ---------------------------------------------------------------------
    def ransac(self, Ya, Yb, samples=None, opts: RansacOpts | None = None, **kw):
        """One pair.  samples (H,k) int32 0-based or None (seeded).  Returns Ransac with mask,
        counts, states (RANSAC_CALC_VER2.m:2-201 / RANSAC_CALC_VER_test.m)."""
        ya, yb = _c(Ya, np.float64), _c(Yb, np.float64)
        N = ya.shape[0]
        o = opts or make_opts(**kw)
        s = None
        if samples is not None:
            s = _c(samples, np.int32)
            o.H, o.k = int(s.shape[0]), int(s.shape[1])
        res = np.zeros(1, RESULT_DTYPE)
        mask = np.zeros(max(N, 1), np.uint8)
        counts = np.zeros(max(o.H, 1), np.int32)
        states = np.zeros(max(o.H, 1), np.int8)
        self._ck(self._lib.pre3_ransac(self._h, _ptr(ya), _ptr(yb), N, C.byref(o), _ptr(s), _ptr(res), _ptr(mask),
                                       _ptr(counts), _ptr(states)))
        return unpack_result(res[0], mask[:N].astype(bool), counts[: o.H], states[: o.H])

    def ransac_batch(self, Ya, Yb, n_corr=None, samples=None, opts: RansacOpts | None = None, want_masks=True, **kw):
        """Ya, Yb (P,Nmax,3); n_corr (P,) or None; samples (P,H,k) or None.  Returns
        (results RESULT_DTYPE (P,), masks (P,Nmax) uint8 | None)."""
        ya, yb = _c(Ya, np.float64), _c(Yb, np.float64)
        P, Nmax = ya.shape[0], ya.shape[1]
        o = opts or make_opts(**kw)
        s = None
        if samples is not None:
            s = _c(samples, np.int32)
            o.H, o.k = int(s.shape[1]), int(s.shape[2])
        nc = None if n_corr is None else _c(n_corr, np.int32)
        res = np.zeros(max(P, 1), RESULT_DTYPE)
        masks = np.zeros((max(P, 1), max(Nmax, 1)), np.uint8) if want_masks else None
        self._ck(self._lib.pre3_ransac_batch(self._h, _ptr(ya), _ptr(yb), _ptr(nc), P, Nmax, C.byref(o), _ptr(s),
                                             _ptr(res), _ptr(masks)))
        return res[:P], (masks[:P, :Nmax] if masks is not None else None)

    # ---- frames -> per-feature 3-D points ---------------------------
    def read_xyz_sr4000_batch(self, sr, opts: FrameOpts | None = None, **kw):
        """sr (F,176,rows) float64 = F column-major rows x 176 sr_data matrices.  Returns x, y, z (F,176,144)
        (= 144 x 176 column-major maps) and max_conf (F,)."""
        sr = _c(sr, np.float64)
        F, cols, rows = sr.shape
        if cols != 176:
            raise ValueError("sr must be (F, 176, rows)")
        o = opts or make_frame_opts(rows=rows, **kw)
        x, y, z = (np.zeros((max(F, 1), 176, 144)) for _ in range(3))
        mc = np.zeros(max(F, 1))
        self._ck(self._lib.pre3_read_xyz_sr4000_batch(self._h, _ptr(sr), F, C.byref(o), _ptr(x), _ptr(y), _ptr(z),
                                                      _ptr(mc)))
        return x[:F], y[:F], z[:F], mc[:F]

    def read_xyz_sr4000_batch_dev(self, sr, opts: FrameOpts, x, y, z, max_conf=None):
        """CUDA tensors: sr (F,176,rows) f64; x, y, z (F,176,144) f64 (or all None: only max_conf); max_conf (F,)."""
        self._ck(self._lib.pre3_read_xyz_sr4000_batch_dev(self._h, _ptr(sr), int(sr.shape[0]), C.byref(opts), _ptr(x),
                                                          _ptr(y), _ptr(z), _ptr(max_conf)))

    def features_xyz_batch(self, sr, frames, desc=None, k_count=None, opts: FrameOpts | None = None, **kw):
        """SIFT_extract_save.m:75-88 for F frames.  sr (F,176,rows); frames (F,K,ld) with the 0-based sift (x, y) in
        the first two entries of a feature; desc (F,K,ND) or None.  Returns a dict: xyz_all (F,K,3), keep (F,K) bool,
        n_keep (F,), idx_remain (F,K), xyz (F,K,3) compacted, frames_out (F,K,ld), desc_out (F,K,ND) | None, n_oob."""
        sr, fr = _c(sr, np.float64), _c(frames, np.float64)
        F, cols, rows = sr.shape
        _, K, ld = fr.shape
        o = opts or make_frame_opts(rows=rows, **kw)
        d = None
        cls, ND = 0, 0
        if desc is not None:
            d = _c(desc)
            if d.dtype not in _CLS:
                raise ValueError("Unsupported numeric class")
            cls, ND = _CLS[d.dtype], int(d.shape[2])
        kc = None if k_count is None else _c(k_count, np.int32)
        Fm, Km = max(F, 1), max(K, 1)
        out = {"xyz_all": np.zeros((Fm, Km, 3)), "keep": np.zeros((Fm, Km), np.uint8), "n_keep": np.zeros(Fm, np.int32),
               "idx_remain": np.zeros((Fm, Km), np.int32), "xyz": np.zeros((Fm, Km, 3)),
               "frames_out": np.zeros((Fm, Km, ld)), "desc_out": None if d is None else np.zeros_like(d)}
        oob = np.zeros(1, np.int32)
        self._ck(self._lib.pre3_features_xyz_batch(self._h, _ptr(sr), F, C.byref(o), _ptr(fr), ld, K, _ptr(kc),
                                                   _ptr(out["xyz_all"]), _ptr(out["keep"]), _ptr(out["n_keep"]),
                                                   _ptr(out["idx_remain"]), _ptr(out["xyz"]), _ptr(d), cls, ND,
                                                   _ptr(out["desc_out"]), _ptr(out["frames_out"]), _ptr(oob)))
        out["keep"] = out["keep"].astype(bool)
        out["n_oob"] = int(oob[0])
        return out

    def features_xyz_batch_dev(self, sr, opts: FrameOpts, frames, xyz=None, n_keep=None, idx_remain=None, xyz_all=None,
                               keep=None, desc_in=None, desc_out=None, frames_out=None, k_count=None, n_oob=None):
        """CUDA tensors (see features_xyz_batch for shapes).  Asynchronous on the context's stream."""
        import torch
        F, K, ld = (int(v) for v in frames.shape)
        cls, ND = 0, 0
        if desc_in is not None:
            cls = {torch.float64: L.CLASS_DOUBLE, torch.float32: L.CLASS_SINGLE, torch.int8: L.CLASS_INT8,
                   torch.uint8: L.CLASS_UINT8}[desc_in.dtype]
            ND = int(desc_in.shape[2])
        self._ck(self._lib.pre3_features_xyz_batch_dev(self._h, _ptr(sr), F, C.byref(opts), _ptr(frames), ld, K,
                                                       _ptr(k_count), _ptr(xyz_all), _ptr(keep), _ptr(n_keep),
                                                       _ptr(idx_remain), _ptr(xyz), _ptr(desc_in), cls, ND,
                                                       _ptr(desc_out), _ptr(frames if frames_out is not None else None),
                                                       _ptr(frames_out), _ptr(n_oob)))

    # ---- code_from_dr_ye variant ---------------------------------------------------------
    def vodometry_dr_ye_batch(self, Ya, Yb, n_corr=None, match=None, samples=None, opts: RansacOpts | None = None,
                              want_masks=True, want_counts=False, **kw):
        """RANSAC part of vodometry_dr_ye.m:147-220 for P match sets.  Ya = pset1, Yb = pset2 (P,Nmax,3);
        match (P,Nmax,2) feature ids or None; samples (P,H,4) draws (0-based) or None (seeded sampler).
        Returns (results RESULT_DTYPE (P,), masks (P,Nmax) | None, stat DR_YE_STAT_DTYPE (P,), counts (P,H) | None)."""
        ya, yb = _c(Ya, np.float64), _c(Yb, np.float64)
        P, Nmax = ya.shape[0], ya.shape[1]
        kw.setdefault("k", 4)
        kw.setdefault("max_iteration", 700)
        kw.setdefault("H", 700)
        o = opts or make_opts(method=L.METHOD_DR_YE, **kw)
        s = None
        if samples is not None:
            s = _c(samples, np.int32)
            if s.ndim != 3 or s.shape[2] != 4:
                raise ValueError("samples must be (P, H, 4)")
            o.H = int(s.shape[1])
        mt = None if match is None else _c(match, np.int32)
        nc = None if n_corr is None else _c(n_corr, np.int32)
        res = np.zeros(max(P, 1), RESULT_DTYPE)
        stat = np.zeros(max(P, 1), DR_YE_STAT_DTYPE)
        masks = np.zeros((max(P, 1), max(Nmax, 1)), np.uint8) if want_masks else None
        counts = np.zeros((max(P, 1), max(int(o.H), 1)), np.int32) if want_counts else None
        self._ck(self._lib.pre3_vodometry_dr_ye_batch(self._h, _ptr(ya), _ptr(yb), _ptr(nc), _ptr(mt), P, Nmax,
                                                      C.byref(o), _ptr(s), _ptr(res), _ptr(masks), _ptr(stat),
                                                      _ptr(counts)))
        return (res[:P], masks[:P, :Nmax] if masks is not None else None, stat[:P],
                counts[:P, :int(o.H)] if counts is not None else None)

    def vodometry_dr_ye_batch_dev(self, Ya, Yb, opts: RansacOpts, res, n_corr=None, match=None, samples=None,
                                  masks=None, stat=None, counts=None, pair_id0=0):
        """CUDA tensors: Ya, Yb (P,Nmax,3) f64; res uint8 (P,240); n_corr int32 (P,) | None; match int32
        (P,Nmax,2) | None; samples int32 (P,H,4) | None; masks uint8 (P,Nmax) | None; stat uint8 (P,32) | None;
        counts int32 (P,H) | None.  Asynchronous on the context's stream."""
        P, Nmax = int(Ya.shape[0]), int(Ya.shape[1])
        self._ck(self._lib.pre3_vodometry_dr_ye_batch_dev(self._h, _ptr(Ya), _ptr(Yb), _ptr(n_corr), _ptr(match), P,
                                                          Nmax, C.byref(opts), _ptr(samples), int(pair_id0),
                                                          _ptr(res), _ptr(masks), _ptr(stat), _ptr(counts)))

    # ---- whole pairs --------------------------------------------------------------------
    def pairs(self, desc1, desc2, xyz1, xyz2, opts: RansacOpts | None = None, pair_id0=0, k1_count=None,
              k2_count=None, want_matches=True, want_masks=True, out=None, **kw):
        """desc (P,K,128) f64/f32/u8/i8, xyz (P,K,3) f64 -- numpy (pinned or pageable) host
        arrays.  SIFT_match_save.m:33-53 for P pairs.  Returns (results (P,), matches (P,K1,2)
        int32 | None, masks (P,K1) uint8 | None); the number of valid matches of pair p is
        results['n_matches'][p].  `out`: optional preallocated (res, matches, masks) tuple."""
        d1, d2 = _c(desc1), _c(desc2)
        if d1.dtype != d2.dtype or d1.dtype not in _CLS:
            raise ValueError("Unsupported numeric class")
        x1, x2 = _c(xyz1, np.float64), _c(xyz2, np.float64)
        P, K1, ND = d1.shape
        K2 = d2.shape[1]
        o = opts or make_opts(**kw)
        k1c = None if k1_count is None else _c(k1_count, np.int32)
        k2c = None if k2_count is None else _c(k2_count, np.int32)
        if out is not None:
            res, matches, masks = out
        else:
            res = np.zeros(max(P, 1), RESULT_DTYPE)
            matches = np.zeros((max(P, 1), max(K1, 1), 2), np.int32) if want_matches else None
            masks = np.zeros((max(P, 1), max(K1, 1)), np.uint8) if want_masks else None
        self._ck(self._lib.pre3_pairs(self._h, _ptr(d1), _ptr(d2), _CLS[d1.dtype], _ptr(x1), _ptr(x2), P, K1, K2, ND,
                                      _ptr(k1c), _ptr(k2c), C.byref(o), int(pair_id0), _ptr(res), _ptr(matches),
                                      _ptr(masks)))
        return res[:P], (matches[:P] if matches is not None else None), (masks[:P] if masks is not None else None)

    def sequence(self, desc, xyz, opts: RansacOpts | None = None, pair_id0=0, k_count=None, want_matches=True,
                 want_masks=True, out=None, **kw):
        """F consecutive frames -> F-1 pairs (frame p, frame p+1).  desc (F,K,128), xyz (F,K,3) host arrays.
        Same returns as pairs() with F-1 entries."""
        d = _c(desc)
        if d.dtype not in _CLS:
            raise ValueError("Unsupported numeric class")
        x = _c(xyz, np.float64)
        F, K, ND = d.shape
        P = max(F - 1, 0)
        o = opts or make_opts(**kw)
        kc = None if k_count is None else _c(k_count, np.int32)
        if out is not None:
            res, matches, masks = out
        else:
            res = np.zeros(max(P, 1), RESULT_DTYPE)
            matches = np.zeros((max(P, 1), max(K, 1), 2), np.int32) if want_matches else None
            masks = np.zeros((max(P, 1), max(K, 1)), np.uint8) if want_masks else None
        self._ck(self._lib.pre3_sequence(self._h, _ptr(d), _CLS[d.dtype], _ptr(x), F, K, ND, _ptr(kc), C.byref(o),
                                         int(pair_id0), _ptr(res), _ptr(matches), _ptr(masks)))
        return res[:P], (matches[:P] if matches is not None else None), (masks[:P] if masks is not None else None)

    def sequence_dev(self, desc, xyz, opts: RansacOpts, res, matches=None, masks=None, pair_id0=0, k_count=None):
        """CUDA tensors: desc (F,K,128), xyz (F,K,3) f64; res uint8 (F-1,240); matches int32 (F-1,K,2) | None;
        masks uint8 (F-1,K) | None."""
        import torch
        cls = {torch.float64: L.CLASS_DOUBLE, torch.float32: L.CLASS_SINGLE, torch.int8: L.CLASS_INT8,
               torch.uint8: L.CLASS_UINT8}[desc.dtype]
        F, K, ND = desc.shape
        for t in (desc, xyz, res):
            assert t.is_cuda and t.is_contiguous()
        self._ck(self._lib.pre3_sequence_dev(self._h, _ptr(desc), cls, _ptr(xyz), F, K, ND, _ptr(k_count),
                                             C.byref(opts), int(pair_id0), _ptr(res), _ptr(matches), _ptr(masks)))

    # ---- config 4: 1-point-RANSAC EKF hypotheses -------------------------------------------
    def ekf_support(self, xi, cam, pattern, z_id, z_euc, threshold):
        """compute_hypothesis_support_fast.m:27-116 for B states.  xi (B,n) [= n x B column-major];
        pattern (n,4) 0/1; z_id (n_id,2), z_euc (n_euc,2) [= 2 x n column-major].
        Returns support (B,) int32, li_id (B,n_id) bool, li_euc (B,n_euc) bool."""
        xi = _c(np.atleast_2d(xi), np.float64)
        B, n = xi.shape
        pat = _c(np.asarray(pattern, np.float64).T)  # column-major n x 4
        if pat.shape != (4, n):
            raise ValueError("state_vector_pattern must be n x 4")
        zi, ze = _c(z_id, np.float64).reshape(-1, 2), _c(z_euc, np.float64).reshape(-1, 2)
        n_id, n_euc = len(zi), len(ze)
        sup = np.zeros(max(B, 1), np.int32)
        li = np.zeros((max(B, 1), max(n_id, 1)), np.uint8)
        le = np.zeros((max(B, 1), max(n_euc, 1)), np.uint8)
        c = make_cam(cam)
        self._ck(self._lib.pre3_ekf_support(self._h, _ptr(xi), n, B, C.byref(c), _ptr(pat), _ptr(zi), n_id, _ptr(ze),
                                            n_euc, float(threshold), _ptr(sup), _ptr(li), _ptr(le)))
        return sup[:B], li.reshape(-1)[:B * n_id].reshape(B, n_id).astype(bool), \
            le.reshape(-1)[:B * n_euc].reshape(B, n_euc).astype(bool)

    def ransac_hypotheses_batch(self, frames: dict, sel=None, opts: EkfOpts | None = None, frame_id0=0,
                                want_supports=False, **kw):
        """ransac_hypotheses.m:27-85 for Fr frames.  `frames`: dict of numpy arrays with a leading
        frame axis -- x (Fr,n), P (Fr,n,n) column-major per frame, type/pos (Fr,F) int32,
        has_z/ic/li0 (Fr,F) uint8, z/h (Fr,F,2), Hcam (Fr,F,13,2), Hfeat (Fr,F,6,2), R (Fr,F,2,2), plus
        scalars std_z and cam (3pre_b200.synth_ekf layout).  sel: (Fr,H,3) int32 or None (seeded).
        Returns (results (Fr,) EKF_RESULT_DTYPE, li_inlier (Fr,F) uint8, supports (Fr,H) | None)."""
        x, P = _c(frames["x"], np.float64), _c(frames["P"], np.float64)
        Fr, n = x.shape
        F = np.asarray(frames["type"]).shape[1]
        o = opts or make_ekf_opts(**kw)
        if sel is not None:
            sel = _c(sel, np.int32)
            if sel.shape != (Fr, o.H, 3):
                raise ValueError("sel must be (Fr, H, 3)")
        ty, po = _c(frames["type"], np.int32), _c(frames["pos"], np.int32)
        hz, ic = _c(frames["has_z"], np.uint8), _c(frames["ic"], np.uint8)
        li = _c(frames["li0"], np.uint8).copy()
        z, h = _c(frames["z"], np.float64), _c(frames["h"], np.float64)
        Hc, Hf, R = _c(frames["Hcam"], np.float64), _c(frames["Hfeat"], np.float64), _c(frames["R"], np.float64)
        res = np.zeros(max(Fr, 1), EKF_RESULT_DTYPE)
        sup = np.zeros((max(Fr, 1), max(o.H, 1)), np.int32) if want_supports else None
        c = make_cam(frames["cam"])
        self._ck(self._lib.pre3_ransac_hypotheses_batch(
            self._h, Fr, n, F, _ptr(x), _ptr(P), float(frames["std_z"]), C.byref(c), _ptr(ty), _ptr(po), _ptr(hz),
            _ptr(ic), _ptr(z), _ptr(h), _ptr(Hc), _ptr(Hf), _ptr(R), _ptr(sel), C.byref(o), int(frame_id0), _ptr(li),
            _ptr(res), _ptr(sup)))
        return res[:Fr], li[:Fr], (sup[:Fr] if sup is not None else None)

    def ransac_hypotheses_batch_dev(self, frames: dict, opts: EkfOpts, li_inlier, res, sel=None, supports=None,
                                    frame_id0=0):
        """Same on CUDA tensors (3pre_b200.synth_ekf.make_ekf_frames(device='cuda') layout);
        res: uint8 (Fr,32); li_inlier: uint8 (Fr,F) in/out.  Stream-ordered, no sync."""
        Fr, n = frames["x"].shape
        F = frames["type"].shape[1]
        c = make_cam(frames["cam"])
        self._ck(self._lib.pre3_ransac_hypotheses_batch_dev(
            self._h, Fr, n, F, _ptr(frames["x"]), _ptr(frames["P"]), float(frames["std_z"]), C.byref(c),
            _ptr(frames["type"]), _ptr(frames["pos"]), _ptr(frames["has_z"]), _ptr(frames["ic"]), _ptr(frames["z"]),
            _ptr(frames["h"]), _ptr(frames["Hcam"]), _ptr(frames["Hfeat"]), _ptr(frames["R"]), _ptr(sel),
            C.byref(opts), int(frame_id0), _ptr(li_inlier), _ptr(res), _ptr(supports)))

    # ---- EKF partial updates (first part) ------------------------------
    def ekf_update_batch(self, frames: dict, sel, r_diag=1.0, x=None, P=None):
        """update.m:27-56 with z, h, H stacked from the features flagged in sel (Fr,F) uint8 and R = r_diag * eye:
        ekf_update_li_inliers.m / ekf_update_hi_inliers.m.  `frames` as in ransac_hypotheses_batch; x / P override
        frames['x'] / frames['P'] (the hi-inlier update runs on x_k_k / p_k_k).  Returns (x_k_k (Fr,n), p_k_k (Fr,n,n)
        column-major per frame, m (Fr,) rows of the stacked system)."""
        x = _c(frames["x"] if x is None else x, np.float64)
        P = _c(frames["P"] if P is None else P, np.float64)
        Fr, n = x.shape
        F = np.asarray(frames["type"]).shape[1]
        sel = _c(sel, np.uint8)
        ty, po = _c(frames["type"], np.int32), _c(frames["pos"], np.int32)
        z, h = _c(frames["z"], np.float64), _c(frames["h"], np.float64)
        Hc, Hf = _c(frames["Hcam"], np.float64), _c(frames["Hfeat"], np.float64)
        xo, Po, m = np.zeros_like(x), np.zeros_like(P), np.zeros(max(Fr, 1), np.int32)
        self._ck(self._lib.pre3_ekf_update_batch(self._h, Fr, n, F, _ptr(x), _ptr(P), _ptr(ty), _ptr(po), _ptr(sel),
                                                 _ptr(z), _ptr(h), _ptr(Hc), _ptr(Hf), float(r_diag), _ptr(xo), _ptr(Po),
                                                 _ptr(m)))
        return xo, Po, m[:Fr]

    def ekf_update_batch_dev(self, frames: dict, sel, x_out, P_out, r_diag=1.0, x=None, P=None, m_out=None):
        """Same on CUDA tensors (make_ekf_frames(device='cuda') layout); stream-ordered, no sync."""
        x = frames["x"] if x is None else x
        P = frames["P"] if P is None else P
        Fr, n = x.shape
        F = frames["type"].shape[1]
        self._ck(self._lib.pre3_ekf_update_batch_dev(self._h, Fr, n, F, _ptr(x), _ptr(P), _ptr(frames["type"]),
                                                     _ptr(frames["pos"]), _ptr(sel), _ptr(frames["z"]),
                                                     _ptr(frames["h"]), _ptr(frames["Hcam"]), _ptr(frames["Hfeat"]),
                                                     float(r_diag), _ptr(x_out), _ptr(P_out), _ptr(m_out)))

    def ekf_update_dense(self, x, P, H, R, z, h):
        """[x_k_k, p_k_k, K] = update(x, p, H, R, z, h) (update.m:27-56) with dense numpy matrices in their mathematical
        orientation: P (n,n), H (m,n), R (m,m).  Returns (x_k_k (n,), p_k_k (n,n), K (n,m))."""
        x = _c(np.asarray(x, np.float64).ravel())
        n = x.size
        z = _c(np.asarray(z, np.float64).ravel())
        h = _c(np.asarray(h, np.float64).ravel())
        m = z.size
        Pc = _c(np.asarray(P, np.float64).T)        # column-major
        Hc = _c(np.asarray(H, np.float64).reshape(m, n).T)
        Rc = _c(np.asarray(R, np.float64).reshape(m, m).T)
        xo, Po, Ko = np.zeros(n), np.zeros((n, n)), np.zeros((max(m, 1), n))
        self._ck(self._lib.pre3_ekf_update_dense(self._h, n, m, _ptr(x), _ptr(Pc), _ptr(Hc), _ptr(Rc), _ptr(z), _ptr(h),
                                                 _ptr(xo), _ptr(Po), _ptr(Ko)))
        return xo, Po.T.copy(), Ko[:m].T.copy()

    def ekf_rescue_hi_inliers_batch_dev(self, frames: dict, P_kk, li, hi, h=None, Hcam=None, Hfeat=None):
        """rescue_hi_inliers.m:35-46 on CUDA tensors: hi (Fr,F) uint8 is written where ic == 1 and li == 0.
        h / Hcam / Hfeat: the measurements re-predicted at x_k_k (default: the frames' own)."""
        Fr, F = frames["type"].shape
        n = frames["x"].shape[1]
        self._ck(self._lib.pre3_ekf_rescue_hi_inliers_batch_dev(
            self._h, Fr, n, F, _ptr(P_kk), _ptr(frames["type"]), _ptr(frames["pos"]), _ptr(frames["ic"]), _ptr(li),
            _ptr(frames["z"]), _ptr(frames["h"] if h is None else h), _ptr(frames["Hcam"] if Hcam is None else Hcam),
            _ptr(frames["Hfeat"] if Hfeat is None else Hfeat), _ptr(hi)))

    def ekf_predict_measurements_batch(self, x, cam, n_rows, n_cols, types, pos, has_h, h_in):
        """rescue_hi_inliers.m:32-33 (predict_camera_measurements + calculate_derivatives at x) for Fr frames.
        x (Fr,n); types, pos (Fr,F) int32; has_h (Fr,F) | None; h_in (Fr,F,2).  Returns h (Fr,F,2), has_h (Fr,F) bool,
        predicted (Fr,F) bool, Hcam (Fr,F,13,2), Hfeat (Fr,F,6,2)."""
        x = _c(x, np.float64)
        ty, ps = _c(types, np.int32), _c(pos, np.int32)
        Fr, n = x.shape
        F = ty.shape[1]
        hh = None if has_h is None else _c(np.asarray(has_h).astype(np.uint8))
        hi = _c(h_in, np.float64)
        h, has, pred = np.zeros((Fr, F, 2)), np.zeros((Fr, F), np.uint8), np.zeros((Fr, F), np.uint8)
        Hc, Hf = np.zeros((Fr, F, 13, 2)), np.zeros((Fr, F, 6, 2))
        c = make_cam(cam)
        self._ck(self._lib.pre3_ekf_predict_measurements_batch(self._h, Fr, n, F, _ptr(x), C.byref(c), int(n_rows),
                                                               int(n_cols), _ptr(ty), _ptr(ps), _ptr(hh), _ptr(hi),
                                                               _ptr(h), _ptr(has), _ptr(pred), _ptr(Hc), _ptr(Hf)))
        return h, has.astype(bool), pred.astype(bool), Hc, Hf

    def ekf_predict_measurements_batch_dev(self, x, cam, n_rows, n_cols, types, pos, has_h, h_in, h_out, has_h_out,
                                           predicted, Hcam, Hfeat):
        """CUDA tensors, shapes as the host form; stream-ordered."""
        Fr, n = x.shape
        c = make_cam(cam)
        self._ck(self._lib.pre3_ekf_predict_measurements_batch_dev(
            self._h, Fr, n, int(types.shape[1]), _ptr(x), C.byref(c), int(n_rows), int(n_cols), _ptr(types), _ptr(pos),
            _ptr(has_h), _ptr(h_in), _ptr(h_out), _ptr(has_h_out), _ptr(predicted), _ptr(Hcam), _ptr(Hfeat)))

    # ---- device-pointer entry points (torch CUDA tensors, stream-ordered, no sync) --------
    def pairs_dev(self, desc1, desc2, xyz1, xyz2, opts: RansacOpts, res, matches=None, masks=None, pair_id0=0,
                  k1_count=None, k2_count=None):
        """CUDA tensors: desc (P,K,128) f64/f32/u8/i8; xyz (P,K,3) f64; res uint8 (P,240);
        matches int32 (P,K1,2) | None; masks uint8 (P,K1) | None."""
        import torch
        cls = {torch.float64: L.CLASS_DOUBLE, torch.float32: L.CLASS_SINGLE, torch.int8: L.CLASS_INT8,
               torch.uint8: L.CLASS_UINT8}[desc1.dtype]
        P, K1, ND = desc1.shape
        K2 = desc2.shape[1]
        for t in (desc1, desc2, xyz1, xyz2, res):
            assert t.is_cuda and t.is_contiguous()
        self._ck(self._lib.pre3_pairs_dev(self._h, _ptr(desc1), _ptr(desc2), cls, _ptr(xyz1), _ptr(xyz2), P, K1, K2,
                                          ND, _ptr(k1_count), _ptr(k2_count), C.byref(opts), int(pair_id0),
                                          _ptr(res), _ptr(matches), _ptr(masks)))

    def siftmatch_batch_dev(self, L1, L2, pairs, score, n_out, thresh=1.5, k1_count=None, k2_count=None):
        import torch
        cls = {torch.float64: L.CLASS_DOUBLE, torch.float32: L.CLASS_SINGLE, torch.int8: L.CLASS_INT8,
               torch.uint8: L.CLASS_UINT8}[L1.dtype]
        P, K1, ND = L1.shape
        K2 = L2.shape[1]
        self._ck(self._lib.pre3_siftmatch_batch_dev(self._h, _ptr(L1), _ptr(L2), cls, P, K1, K2, ND, _ptr(k1_count),
                                                    _ptr(k2_count), float(thresh), _ptr(pairs), _ptr(score),
                                                    _ptr(n_out)))

    def cov_est_ransac_batch(self, Ya, Yb, R, T, n_corr=None, masks=None):
        """cov_est_RANSAC_deriv.m for P pairs.  Ya, Yb (P,Nmax,3); R (P,3,3) with Ya ~ R Yb + T; T (P,3); masks (P,Nmax)
        = the winner's support set or None.  Returns a list of dicts: cov (7,7) over [T1 T2 T3 q1 q2 q3 q4], G2tot
        (7,7), Gtot (7,), dA_dz (7,6), Etot, s2, n, status."""
        ya, yb = _c(Ya, np.float64), _c(Yb, np.float64)
        P, Nmax = ya.shape[0], ya.shape[1]
        Rc = _c(np.asarray(R, np.float64).reshape(P, 3, 3).transpose(0, 2, 1)).reshape(P, 9)
        Tc = _c(np.asarray(T, np.float64).reshape(P, 3))
        nc = None if n_corr is None else _c(n_corr, np.int32)
        mk = None if masks is None else _c(np.asarray(masks).astype(np.uint8))
        out = np.zeros(max(P, 1), COV_RESULT_DTYPE)
        self._ck(self._lib.pre3_cov_est_ransac_batch(self._h, _ptr(ya), _ptr(yb), _ptr(nc), _ptr(mk), P, Nmax, _ptr(Rc),
                                                     _ptr(Tc), _ptr(out)))
        return [unpack_cov(out[p]) for p in range(P)]

    def cov_est_ransac_batch_dev(self, Ya, Yb, RT, rt_stride, out, n_corr=None, masks=None):
        """CUDA tensors.  RT: tensor (or data pointer holder) whose element p * rt_stride starts R (9, column-major)
        then T (3); out: (P, COV_RESULT_DTYPE.itemsize) uint8."""
        self._ck(self._lib.pre3_cov_est_ransac_batch_dev(self._h, _ptr(Ya), _ptr(Yb), _ptr(n_corr), _ptr(masks),
                                                         int(Ya.shape[0]), int(Ya.shape[1]), _ptr(RT), int(rt_stride),
                                                         _ptr(out)))

    def ransac_batch_dev(self, Ya, Yb, opts: RansacOpts, res, n_corr=None, samples=None, masks=None):
        P, Nmax = Ya.shape[0], Ya.shape[1]
        self._ck(self._lib.pre3_ransac_batch_dev(self._h, _ptr(Ya), _ptr(Yb), _ptr(n_corr), P, Nmax, C.byref(opts),
                                                 _ptr(samples), _ptr(res), _ptr(masks)))

    def distance_threshold_dev(self, Yb, out):
        self._ck(self._lib.pre3_distance_threshold_dev(self._h, _ptr(Yb), Yb.shape[0], _ptr(out)))

    def ransac_block_dev(self, Ya, Yb, opts: RansacOpts, h0: int, Hloc: int, thr: float, key, errsum=None,
                         samples=None):
        """Local best of hypotheses [h0, h0+Hloc) of one pair: key (1,) int64/uint64 tensor =
        (count << 32) | (0xFFFFFFFF - global id); errsum (1,) f64 tensor | None."""
        self._ck(self._lib.pre3_ransac_block_dev(self._h, _ptr(Ya), _ptr(Yb), Ya.shape[0], C.byref(opts),
                                                 _ptr(samples), int(h0), int(Hloc), float(thr), _ptr(key),
                                                 _ptr(errsum)))

    def ransac_block_select_dev(self, Ya, Yb, opts: RansacOpts, h0: int, Hloc: int, thr: float, res, mask=None,
                                samples=None):
        """Reference-exact local winner of hypotheses [h0, h0+Hloc): res uint8 (240,) tensor
        (best_sample is local: add h0), mask uint8 (N,) | None."""
        self._ck(self._lib.pre3_ransac_block_select_dev(self._h, _ptr(Ya), _ptr(Yb), Ya.shape[0], C.byref(opts),
                                                        _ptr(samples), int(h0), int(Hloc), float(thr), _ptr(res),
                                                        _ptr(mask)))

    def ransac_split_local_dev(self, Ya, Yb, opts: RansacOpts, h0: int, Hloc: int, mode: int, key, res=None, mask=None,
                               samples=None):
        """Stream-ordered hypothesis-block split, local part (pre3_ransac_split_local_dev): key int64 tensor (1,)
        (mode 0) or (2,) (mode 1); res uint8 (240,), mask uint8 (N,) for mode 1."""
        self._ck(self._lib.pre3_ransac_split_local_dev(self._h, _ptr(Ya), _ptr(Yb), Ya.shape[0], C.byref(opts),
                                                       _ptr(samples), int(h0), int(Hloc), int(mode), _ptr(key),
                                                       _ptr(res), _ptr(mask)))

    def ransac_split_finish_dev(self, Ya, Yb, opts: RansacOpts, h0: int, Hloc: int, mode: int, exchanged, world: int,
                                rank: int, res, mask=None, samples=None):
        self._ck(self._lib.pre3_ransac_split_finish_dev(self._h, _ptr(Ya), _ptr(Yb), Ya.shape[0], C.byref(opts),
                                                        _ptr(samples), int(h0), int(Hloc), int(mode), _ptr(exchanged),
                                                        int(world), int(rank), _ptr(res), _ptr(mask)))

    def ransac_finish_dev(self, Ya, Yb, opts: RansacOpts, winner_id: int, thr: float, res, mask=None,
                          sample_of_winner=None):
        self._ck(self._lib.pre3_ransac_finish_dev(self._h, _ptr(Ya), _ptr(Yb), Ya.shape[0], C.byref(opts),
                                                  _ptr(sample_of_winner), int(winner_id), float(thr), _ptr(res),
                                                  _ptr(mask)))


def R2q(R):
    """slamToolbox R2q (R2q.m:11-55): q = [a -b -c -d]'.  R is a 3x3 array."""
    Rc = _c(np.asarray(R, np.float64).T).reshape(9)  # column-major bytes
    q = np.zeros(4)
    L.load().pre3_R2q(_ptr(Rc), _ptr(q))
    return q

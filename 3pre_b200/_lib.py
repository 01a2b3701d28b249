"""ctypes binding of libpre3.so (include/pre3.h).  No compute lives on this side.

The library is the product: if it is missing this module raises -- there is no CPU or
PyTorch fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libpre3.so")

OK, ERR_ARG, ERR_CUDA, ERR_ALLOC, ERR_CLASS = 0, -1, -2, -3, -4
CLASS_DOUBLE, CLASS_SINGLE, CLASS_INT8, CLASS_UINT8 = 0, 1, 2, 3
METHOD_SVD, METHOD_HORN, METHOD_DR_YE = 0, 1, 2
MATCH_AUTO, MATCH_EXACT, MATCH_TC = 0, 1, 2
TIMING_NCAT = 15


class RansacOpts(C.Structure):
    """pre3_ransac_opts"""
    _fields_ = [
        ("method", C.c_int32),
        ("k", C.c_int32),
        ("max_iteration", C.c_int32),
        ("adaptive", C.c_int32),
        ("H", C.c_int32),
        ("reserved", C.c_int32),
        ("distance_threshold", C.c_double),
        ("ratio", C.c_double),
        ("seed", C.c_uint64),
    ]


class PairResult(C.Structure):
    """pre3_pair_result (240 bytes)"""
    _fields_ = [
        ("status", C.c_int32),
        ("state", C.c_int32),
        ("best_fit", C.c_int32),
        ("best_sample", C.c_int32),
        ("best_iter", C.c_int32),
        ("n_iter", C.c_int32),
        ("n_consumed", C.c_int32),
        ("n_matches", C.c_int32),
        ("thr", C.c_double),
        ("error_sum", C.c_double),
        ("R", C.c_double * 9),
        ("T", C.c_double * 3),
        ("R_hyp", C.c_double * 9),
        ("T_hyp", C.c_double * 3),
    ]


class DrYeStat(C.Structure):
    """pre3_dr_ye_stat (32 bytes)"""
    _fields_ = [("error_mean", C.c_double), ("error_std", C.c_double), ("dist", C.c_double),
                ("n_iteration_ransac", C.c_int32), ("n_loops", C.c_int32)]


class FrameOpts(C.Structure):
    """pre3_frame_opts (24 bytes)"""
    _fields_ = [("sigma", C.c_double), ("boundary", C.c_int32), ("mode", C.c_int32), ("rows", C.c_int32),
                ("use_confidence", C.c_int32)]


class Cam(C.Structure):
    """pre3_cam"""
    _fields_ = [("f", C.c_double), ("Cx", C.c_double), ("Cy", C.c_double), ("k1", C.c_double), ("k2", C.c_double)]


class EkfOpts(C.Structure):
    """pre3_ekf_opts"""
    _fields_ = [("n_hyp_init", C.c_int32), ("H", C.c_int32), ("adaptive", C.c_int32), ("reserved", C.c_int32),
                ("seed", C.c_uint64)]


class EkfResult(C.Structure):
    """pre3_ekf_result (32 bytes)"""
    _fields_ = [("status", C.c_int32), ("n_evaluated", C.c_int32), ("best_hyp", C.c_int32),
                ("max_support", C.c_int32), ("num_ic", C.c_int32), ("m", C.c_int32), ("n_hyp", C.c_double)]


# name -> (restype, argtypes); every symbol include/pre3.h declares
_VP, _I, _D, _I64, _U32, _U64 = C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_uint32, C.c_uint64
_OPTS = C.POINTER(RansacOpts)
SYMBOLS = {
    "pre3_create": (_I, [C.POINTER(_VP), _I]),
    "pre3_destroy": (None, [_VP]),
    "pre3_last_error": (C.c_char_p, [_VP]),
    "pre3_version": (C.c_char_p, []),
    "pre3_set_stream": (_I, [_VP, _VP]),
    "pre3_set_match_engine": (_I, [_VP, _I]),
    "pre3_sync": (_I, [_VP]),
    "pre3_launch_count": (_I64, [_VP]),
    "pre3_set_graphs": (_I, [_VP, _I]),
    "pre3_set_pipeline": (_I, [_VP, _I]),
    "pre3_transfer_bytes": (_I, [_VP, _VP, _VP]),
    "pre3_eval_schedule": (_I, [_OPTS, _VP, _I]),
    "pre3_eval_schedule_for": (_I, [_OPTS, _I, _VP, _I]),
    "pre3_timing_enable": (_I, [_VP, _I]),
    "pre3_timing_read": (_I, [_VP, _VP, _VP]),
    "pre3_timing_name": (C.c_char_p, [_I]),
    "pre3_measure_fp32_peak": (_I, [_VP, _VP]),
    "pre3_measure_fp32_peak_mode": (_I, [_VP, _I, _VP]),
    "pre3_siftmatch": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _D, _VP, _VP, _VP]),
    "pre3_siftmatch_batch": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, _VP, _VP, _D, _VP, _VP, _VP]),
    "pre3_siftmatch_batch_dev": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, _VP, _VP, _D, _VP, _VP, _VP]),
    "pre3_cov_est_ransac_batch": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, _VP, _VP, _VP]),
    "pre3_cov_est_ransac_batch_dev": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, _VP, _I, _VP]),
    "pre3_ekf_predict_measurements_batch": (_I, [_VP, _I, _I, _I, _VP, _VP, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP,
                                                 _VP, _VP]),
    "pre3_ekf_predict_measurements_batch_dev": (_I, [_VP, _I, _I, _I, _VP, _VP, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP,
                                                     _VP, _VP, _VP]),
    "pre3_siftmatch_sweep": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, _VP, _D, _VP, _VP, _VP]),
    "pre3_siftmatch_sweep_dev": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, _VP, _D, _VP, _VP, _VP]),
    "pre3_matching_sift_based_batch": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _D, _VP, _VP,
                                            _VP, _VP, _VP]),
    "pre3_find_transform_matrix": (_I, [_VP, _VP, _VP, _I, _VP, _VP, _VP]),
    "pre3_horn": (_I, [_VP, _VP, _VP, _I, _I, _VP, _VP, _VP, _VP]),
    "pre3_fit_batch": (_I, [_VP, _VP, _VP, _I, _VP, _I, _I, _I, _VP, _VP, _VP]),
    "pre3_score_batch": (_I, [_VP, _VP, _VP, _I, _VP, _VP, _I, _D, _VP, _VP, _VP]),
    "pre3_ransac": (_I, [_VP, _VP, _VP, _I, _OPTS, _VP, _VP, _VP, _VP, _VP]),
    "pre3_ransac_batch": (_I, [_VP, _VP, _VP, _VP, _I, _I, _OPTS, _VP, _VP, _VP]),
    "pre3_ransac_batch_dev": (_I, [_VP, _VP, _VP, _VP, _I, _I, _OPTS, _VP, _VP, _VP]),
    "pre3_ekf_update_batch": (_I, [_VP, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _D, _VP, _VP, _VP]),
    "pre3_ekf_update_batch_dev": (_I, [_VP, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _D, _VP, _VP, _VP]),
    "pre3_ekf_update_dense": (_I, [_VP, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "pre3_ekf_update_dense_dev": (_I, [_VP, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "pre3_ekf_rescue_hi_inliers_batch_dev": (_I, [_VP, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "pre3_read_xyz_sr4000_batch": (_I, [_VP, _VP, _I, C.POINTER(FrameOpts), _VP, _VP, _VP, _VP]),
    "pre3_read_xyz_sr4000_batch_dev": (_I, [_VP, _VP, _I, C.POINTER(FrameOpts), _VP, _VP, _VP, _VP]),
    "pre3_features_xyz_batch": (_I, [_VP, _VP, _I, C.POINTER(FrameOpts), _VP, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP,
                                     _VP, _I, _I, _VP, _VP, _VP]),
    "pre3_features_xyz_batch_dev": (_I, [_VP, _VP, _I, C.POINTER(FrameOpts), _VP, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP,
                                         _VP, _I, _I, _VP, _VP, _VP, _VP]),
    "pre3_vodometry_dr_ye_batch": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, _OPTS, _VP, _VP, _VP, _VP, _VP]),
    "pre3_vodometry_dr_ye_batch_dev": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, _OPTS, _VP, _U32, _VP, _VP, _VP, _VP]),
    "pre3_pairs": (_I, [_VP, _VP, _VP, _I, _VP, _VP, _I, _I, _I, _I, _VP, _VP, _OPTS, _U32, _VP, _VP, _VP]),
    "pre3_pairs_dev": (_I, [_VP, _VP, _VP, _I, _VP, _VP, _I, _I, _I, _I, _VP, _VP, _OPTS, _U32, _VP, _VP, _VP]),
    "pre3_sequence": (_I, [_VP, _VP, _I, _VP, _I, _I, _I, _VP, _OPTS, _U32, _VP, _VP, _VP]),
    "pre3_sequence_dev": (_I, [_VP, _VP, _I, _VP, _I, _I, _I, _VP, _OPTS, _U32, _VP, _VP, _VP]),
    "pre3_ransac_block_dev": (_I, [_VP, _VP, _VP, _I, _OPTS, _VP, _I64, _I, _D, _VP, _VP]),
    "pre3_ransac_block_select_dev": (_I, [_VP, _VP, _VP, _I, _OPTS, _VP, _I64, _I, _D, _VP, _VP]),
    "pre3_ransac_finish_dev": (_I, [_VP, _VP, _VP, _I, _OPTS, _VP, _I64, _D, _VP, _VP]),
    "pre3_distance_threshold_dev": (_I, [_VP, _VP, _I, _VP]),
    "pre3_ransac_split_local_dev": (_I, [_VP, _VP, _VP, _I, _OPTS, _VP, _I64, _I, _I, _VP, _VP, _VP]),
    "pre3_ransac_split_finish_dev": (_I, [_VP, _VP, _VP, _I, _OPTS, _VP, _I64, _I, _I, _VP, _I, _I, _VP, _VP]),
    "pre3_R2q": (None, [_VP, _VP]),
    "pre3_ekf_support": (_I, [_VP, _VP, _I, _I, C.POINTER(Cam), _VP, _VP, _I, _VP, _I, _D, _VP, _VP, _VP]),
    "pre3_ransac_hypotheses_batch": (_I, [_VP, _I, _I, _I, _VP, _VP, _D, C.POINTER(Cam), _VP, _VP, _VP, _VP, _VP, _VP,
                                          _VP, _VP, _VP, _VP, C.POINTER(EkfOpts), _U32, _VP, _VP, _VP]),
    "pre3_ransac_hypotheses_batch_dev": (_I, [_VP, _I, _I, _I, _VP, _VP, _D, C.POINTER(Cam), _VP, _VP, _VP, _VP, _VP,
                                              _VP, _VP, _VP, _VP, _VP, C.POINTER(EkfOpts), _U32, _VP, _VP, _VP]),
    "pre3_ekf_eval_schedule": (_I, [C.POINTER(EkfOpts), _VP, _I]),
    "pre3_measure_fp64_peak": (_I, [_VP, _VP]),
    "pre3_measure_tmem_read": (_I, [_VP, _VP]),
}

_lib = None


class Pre3Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pre3 error {code}: {msg}")
        self.code = code


def load():
    """dlopen libpre3.so and type every entry point.  Raises if the CUDA library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python 3pre_b200/build.py` "
            "(__graft_entry__.build()).  3pre_b200 has no CPU / PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib

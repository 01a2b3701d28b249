"""pre3-b200: the B200 (sm_100a) drop-in for the frame-to-frame motion-estimation hot path of
ahtamjidi/3PRE (SIFT descriptor matching -> minimal-sample rigid fits -> hypothesis support
-> selection + refit, and the 1-point-RANSAC EKF hypothesis support).

The product is 3pre_b200/lib/libpre3.so (CUDA, C ABI in include/pre3.h).  This package is
its host-side mirror of the reference's MATLAB interface; it contains no compute and no
fallback: without the built library (or without a B200) every call raises.
"""
from . import _lib
from ._lib import Pre3Error, RansacOpts, PairResult, Cam, EkfOpts, EkfResult, DrYeStat, FrameOpts
from .api import (Context, R2q, make_opts, unpack_result, RESULT_DTYPE, make_cam, make_ekf_opts,
                  EKF_RESULT_DTYPE, DR_YE_STAT_DTYPE, make_frame_opts, COV_RESULT_DTYPE, unpack_cov)

__all__ = ["Context", "Pre3Error", "RansacOpts", "PairResult", "R2q", "make_opts", "unpack_result", "RESULT_DTYPE", "COV_RESULT_DTYPE", "unpack_cov",
           "Cam", "EkfOpts", "EkfResult", "make_cam", "make_ekf_opts", "EKF_RESULT_DTYPE",
           "DrYeStat", "DR_YE_STAT_DTYPE", "FrameOpts", "make_frame_opts"]

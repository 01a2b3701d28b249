"""Drive the REFERENCE's own siftmatch MEX gateway (matlab_code/sift/siftmatch.c:139-250)
compiled into oracle/_ref/libsiftmatch_ref.so by oracle/Makefile.

TEST INFRASTRUCTURE ONLY.  The .so is built from the sources where they lie under
/root/reference (never copied); on the GPU box the prebuilt .so that travelled with
the snapshot is used.  `available()` tells whether it exists.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libsiftmatch_ref.so")

# mxClassID values of oracle/mex_stub/mex.h
_CLS = {np.dtype(np.float64): 6, np.dtype(np.float32): 7, np.dtype(np.int8): 8, np.dtype(np.uint8): 9,
        np.dtype(np.int32): 12}


class _MxArray(C.Structure):
    _fields_ = [("cls", C.c_int), ("m", C.c_size_t), ("n", C.c_size_t), ("ndim", C.c_int),
                ("is_complex", C.c_int), ("owns_data", C.c_int), ("data", C.c_void_p)]


_libs = {}


def available() -> bool:
    if not os.path.exists(_SO) and os.path.exists("/root/reference/matlab_code/sift/siftmatch.c"):
        from . import oracle as _o
        _o.build()
    return os.path.exists(_SO)


def lib(so=None):
    """The reference gateway (default) or any other siftmatch MEX gateway linked against the
    same stub runtime (tests build the product's gateway that way to drive both identically)."""
    path = so or _SO
    if path not in _libs:
        if so is None and not available():
            raise RuntimeError("oracle/_ref/libsiftmatch_ref.so not built (no /root/reference here?)")
        L = C.CDLL(path)
        L.stub_wrap.restype = C.POINTER(_MxArray)
        L.stub_wrap.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_void_p]
        L.stub_call_mex.argtypes = [C.c_int, C.POINTER(C.POINTER(_MxArray)), C.c_int, C.POINTER(C.POINTER(_MxArray))]
        L.stub_last_error.restype = C.c_char_p
        L.mxDestroyArray.argtypes = [C.POINTER(_MxArray)]
        _libs[path] = L
    return _libs[path]


class MexError(RuntimeError):
    pass


def siftmatch(L1, L2, thresh=None, nout=2, extra_args=0, so=None):
    """Call the reference gateway: matches = siftmatch(L1, L2[, thresh]).
    L1:(K1,ND), L2:(K2,ND) numpy (== ND x K column-major).  Returns
    (matches (2,n) float64 1-BASED exactly as MATLAB sees it, D (n,) or None)."""
    L = lib(so)
    L1 = np.ascontiguousarray(L1)
    L2 = np.ascontiguousarray(L2)
    keep = [L1, L2]
    ins = [L.stub_wrap(_CLS.get(L1.dtype, 0), L1.shape[1], L1.shape[0], L1.ctypes.data),
           L.stub_wrap(_CLS.get(L2.dtype, 0), L2.shape[1], L2.shape[0], L2.ctypes.data)]
    if thresh is not None:
        t = np.array([float(thresh)])
        keep.append(t)
        ins.append(L.stub_wrap(6, 1, 1, t.ctypes.data))
    for _ in range(extra_args):
        t = np.array([0.0])
        keep.append(t)
        ins.append(L.stub_wrap(6, 1, 1, t.ctypes.data))
    in_arr = (C.POINTER(_MxArray) * len(ins))(*ins)
    out_arr = (C.POINTER(_MxArray) * max(nout, 1))()
    rc = L.stub_call_mex(nout, out_arr, len(ins), in_arr)
    for a in ins:
        L.mxDestroyArray(a)
    if rc != 0:
        raise MexError(L.stub_last_error().decode())
    m = out_arr[0].contents
    n = m.n
    matches = np.ctypeslib.as_array(C.cast(m.data, C.POINTER(C.c_double)), shape=(n, 2)).copy().T if n else np.zeros((2, 0))
    D = None
    if nout > 1:
        d = out_arr[1].contents
        D = np.ctypeslib.as_array(C.cast(d.data, C.POINTER(C.c_double)), shape=(d.n,)).copy() if d.n else np.zeros(0)
        L.mxDestroyArray(out_arr[1])
    L.mxDestroyArray(out_arr[0])
    return matches, D

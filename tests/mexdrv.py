"""Drives the product's MEX gateways (3pre_b200/mex_files/*.cpp) through their real mexFunction, linked against
the same stub MEX runtime that drives the reference's own siftmatch.c gateway (oracle/mex_stub).  Test
infrastructure: MATLAB / Octave do not exist in this image, so this is how `options` structs, `cam` structs,
logical outputs and struct outputs of the gateways are exercised.

    gw = Gateway("RANSAC_CALC_VER2_mex")
    R, T, err, best_fit, state = gw(Ya, Yb, {"DistanceThreshold": 0.05, "MaxIteration": 2000}, nout=5)

Arguments are MATLAB-shaped numpy arrays (rows, cols) -- passed column-major -- scalars, or dicts (1 x 1 structs).
Outputs come back MATLAB-shaped; structs as dicts; logical arrays as bool."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle import refmex

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CLS = {np.dtype(np.float64): 6, np.dtype(np.float32): 7, np.dtype(np.int8): 8, np.dtype(np.uint8): 9,
        np.dtype(np.int32): 12, np.dtype(np.bool_): 3}
_NP = {6: np.float64, 7: np.float32, 8: np.int8, 9: np.uint8, 12: np.int32, 3: np.uint8}


def build_gateway(name: str) -> str:
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, f"libpre3_{name}_gw.so")
    libdir = os.path.join(ROOT, "3pre_b200", "lib")
    src = os.path.join(ROOT, "3pre_b200", "mex_files", name + ".cpp")
    stub = os.path.join(ROOT, "oracle", "mex_stub", "mex_stub.c")
    deps = [src, stub, os.path.join(ROOT, "oracle", "mex_stub", "mex.h"),
            os.path.join(ROOT, "3pre_b200", "mex_files", "pre3_mex_common.h"), os.path.join(ROOT, "include", "pre3.h")]
    if os.path.exists(so) and all(os.path.getmtime(so) > os.path.getmtime(d) for d in deps):
        return so
    subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-o", so, src, "-x", "c", stub, "-x", "none",
                    "-I", os.path.join(ROOT, "oracle", "mex_stub"), "-I", os.path.join(ROOT, "3pre_b200", "mex_files"),
                    "-I", os.path.join(ROOT, "include"), "-L", libdir, "-lpre3", f"-Wl,-rpath,{libdir}"], check=True)
    return so


class Gateway:
    def __init__(self, name: str):
        self.L = refmex.lib(build_gateway(name))
        L = self.L
        L.mxCreateStructMatrix.restype = C.POINTER(refmex._MxArray)
        L.mxCreateStructMatrix.argtypes = [C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_char_p)]
        L.mxSetField.argtypes = [C.POINTER(refmex._MxArray), C.c_size_t, C.c_char_p, C.POINTER(refmex._MxArray)]
        L.mxGetField.restype = C.POINTER(refmex._MxArray)
        L.mxGetField.argtypes = [C.POINTER(refmex._MxArray), C.c_size_t, C.c_char_p]
        L.stub_struct_nfields.argtypes = [C.POINTER(refmex._MxArray)]
        L.stub_struct_field_name.restype = C.c_char_p
        L.stub_struct_field_name.argtypes = [C.POINTER(refmex._MxArray), C.c_int]
        self._keep = []

    def _wrap(self, a):
        L = self.L
        if isinstance(a, dict):
            names = [k.encode() for k in a]
            arr = (C.c_char_p * len(names))(*names)
            self._keep += [names, arr]
            s = L.mxCreateStructMatrix(1, 1, len(names), arr)
            for k, v in a.items():
                L.mxSetField(s, 0, k.encode(), self._wrap(v))  # the struct owns its fields
            return s
        a = np.asarray(a)
        if a.dtype not in _CLS:
            a = a.astype(np.float64)
        if a.ndim == 0:
            a = a.reshape(1, 1)
        elif a.ndim == 1:
            a = a.reshape(1, -1)           # MATLAB row vector
        elif a.ndim > 2:
            a = a.reshape(a.shape[0], -1, order="F")
        f = np.asfortranarray(a)
        self._keep.append(f)
        return L.stub_wrap(_CLS[f.dtype], f.shape[0], f.shape[1], f.ctypes.data)

    def _read(self, p):
        L = self.L
        m = p.contents
        if m.cls == 2:  # struct
            out = {}
            for k in range(L.stub_struct_nfields(p)):
                name = L.stub_struct_field_name(p, k)
                out[name.decode()] = self._read(L.mxGetField(p, 0, name))
            return out
        dt = _NP[m.cls]
        if m.m * m.n == 0:
            return np.zeros((m.m, m.n), dtype=bool if m.cls == 3 else dt)
        flat = np.ctypeslib.as_array(C.cast(m.data, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(m.m * m.n,)).copy()
        arr = flat.reshape((m.m, m.n), order="F")
        return arr.astype(bool) if m.cls == 3 else arr

    def __call__(self, *args, nout=1):
        L = self.L
        self._keep = []
        ins = [self._wrap(a) for a in args]
        in_arr = (C.POINTER(refmex._MxArray) * max(len(ins), 1))(*ins)
        out_arr = (C.POINTER(refmex._MxArray) * max(nout, 1))()
        rc = L.stub_call_mex(nout, out_arr, len(ins), in_arr)
        for a in ins:
            L.mxDestroyArray(a)
        if rc != 0:
            raise refmex.MexError(L.stub_last_error().decode())
        res = []
        for i in range(max(nout, 1)):
            res.append(self._read(out_arr[i]))
            L.mxDestroyArray(out_arr[i])
        self._keep = []
        return res[0] if nout <= 1 else res

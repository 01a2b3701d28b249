"""Builds 3pre_b200/lib/libpre3.so (the C-ABI CUDA library, include/pre3.h) and the MEX
gateways, in-tree, for sm_100a only.

nvcc cross-compiles without a GPU.  -fmad=false: the fp64 fit / rescore code must not be
contracted into FMAs (it shares its operation order with the CPU checker); kernels that want
fused multiply-adds call fma()/fmaf() explicitly.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libpre3.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    """.cu: CUDA translation units; .cpp: host-only helpers (g++ through nvcc, no device code)."""
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _deps():
    out = sources()
    out += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    out.append(os.path.join(ROOT, "include", "pre3.h"))
    return out


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build_lib(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ and link libpre3.so.  Objects are cached per source."""
    if not force and not stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    hdr_t = max(os.path.getmtime(d) for d in _deps() if not d.endswith(".cu"))
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.splitext(os.path.basename(src))[0] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            continue
        if src.endswith(".cpp"):
            cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-c", src, "-o", obj]
        else:
            cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
        if verbose and not src.endswith(".cpp"):
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {src}:\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("libpre3 build failed")
    tmp = LIB + ".tmp"  # link beside the target and swap: a snapshot taken mid-build never sees a half-written .so
    link = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC", "-cudart", "static", "-lpthread"]
    subprocess.run(link, check=True)
    os.replace(tmp, LIB)
    return LIB


MEXDIR = os.path.join(HERE, "mex_files")


def build_mex(force: bool = False):
    """Compile the MEX gateways against the declarations-only shim (object files: the real
    link against libmex / liboctinterp happens on the user's machine)."""
    outdir = os.path.join(LIBDIR, "mex_obj")
    os.makedirs(outdir, exist_ok=True)
    objs = []
    for f in sorted(os.listdir(MEXDIR)):
        if not f.endswith(".cpp"):
            continue
        src, obj = os.path.join(MEXDIR, f), os.path.join(outdir, f[:-4] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src):
            continue
        subprocess.run(["g++", "-O2", "-fPIC", "-Wall", "-c", src, "-o", obj, "-I", os.path.join(MEXDIR, "mex_shim"),
                        "-I", MEXDIR, "-I", os.path.join(ROOT, "include")], check=True)
    return objs


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_mex(force="--force" in sys.argv))

"""CPU suite, part 2: the drop-in boundary (no compute calls -- there is no GPU here).

* libpre3.so loads and exports every symbol include/pre3.h declares;
* struct layouts of the ctypes mirror match the header;
* without a CUDA device the product fails loudly (no CPU fallback);
* the MATLAB-shaped mirror raises the reference's argument errors before any device work.
"""
import ctypes as C
import importlib
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pre3.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"PRE3_API\s+[\w\s\*]+?\b(pre3_\w+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ["pre3_siftmatch", "pre3_ransac", "pre3_pairs", "pre3_horn", "pre3_find_transform_matrix",
              "pre3_score_batch", "pre3_ransac_block_dev", "pre3_create", "pre3_destroy"]:
        assert s in syms


def test_library_exports_every_declared_symbol(pre3):
    lib = pre3._lib.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"libpre3.so does not export {s}"
    # and the ctypes table covers the header one-to-one
    assert sorted(pre3._lib.SYMBOLS) == syms
    assert b"sm_100a" in lib.pre3_version()


def test_every_entry_point_cites_the_reference_interface_it_replaces():
    """include/pre3.h: the comment block in front of every PRE3_API declaration (or of the group it belongs to) names
    a reference file (M/...: .m / .c) with line numbers, or is plumbing (context, timing, measurement helpers)."""
    src = open(HEADER).read()
    plumbing = ("create", "destroy", "last_error", "version", "set_stream", "set_match_engine", "sync", "launch_count",
                "transfer_bytes", "timing", "measure", "eval_schedule")
    # split at blank lines that precede a comment opener: one chunk = comment block + its declarations
    chunks = re.split(r"\n\s*\n(?=/\*)", src)
    cited = set()
    for ch in chunks:
        names = re.findall(r"PRE3_API\s+[\w\s\*]+?\b(pre3_\w+)\s*\(", ch)
        if names and re.search(r"\.(m|c|prj)\b[^\n]*?:\d+|\.(m|c):\d+", ch):
            cited.update(names)
    missing = [s for s in declared_symbols() if s not in cited and not any(k in s for k in plumbing)]
    assert not missing, missing


def test_struct_layouts_of_the_later_additions(pre3):
    L = pre3._lib
    assert C.sizeof(L.DrYeStat) == 32 and pre3.DR_YE_STAT_DTYPE.itemsize == 32
    assert C.sizeof(L.FrameOpts) == 24 and L.FrameOpts.rows.offset == 16
    assert L.TIMING_NCAT == int(re.search(r"#define PRE3_TIMING_NCAT (\d+)", open(HEADER).read()).group(1))


def test_library_is_sm100a_only():
    so = os.path.join(ROOT, "3pre_b200", "lib", "libpre3.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layouts(pre3):
    assert C.sizeof(pre3.PairResult) == 240 and pre3.RESULT_DTYPE.itemsize == 240
    assert C.sizeof(pre3.RansacOpts) == 48
    assert pre3.PairResult.thr.offset == 32 and pre3.PairResult.R.offset == 48
    assert pre3.RESULT_DTYPE.fields["R_hyp"][1] == pre3.PairResult.R_hyp.offset


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_gpu_fails_loudly(pre3):
    with pytest.raises(pre3.Pre3Error) as e:
        pre3.Context()
    assert e.value.code == pre3._lib.ERR_CUDA and "no CPU fallback" in str(e.value)
    m = importlib.import_module("3pre_b200.matlab")
    with pytest.raises(pre3.Pre3Error):
        m.siftmatch(np.zeros((128, 4)), np.zeros((128, 5)))


def test_mirror_argument_errors_match_reference():
    m = importlib.import_module("3pre_b200.matlab")
    a = np.zeros((3, 4))
    with pytest.raises(m.MexError, match="same number of rows"):
        m.siftmatch(a, np.zeros((2, 4)))
    with pytest.raises(m.MexError, match="same class"):
        m.siftmatch(a, a.astype(np.float32))
    with pytest.raises(m.MexError, match="Unsupported numeric class"):
        m.siftmatch(a.astype(np.int32), a.astype(np.int32))
    with pytest.raises(m.MexError, match="Too many output"):
        m.siftmatch(a, a, nargout=3)
    with pytest.raises(m.MexError, match="real scalar"):
        m.siftmatch(a, a, np.array([1.0, 2.0]))
    with pytest.raises(m.MexError, match="at least 4"):
        m.absoluteOrientationQuaternion(np.zeros((3, 3)), np.zeros((3, 3)))
    with pytest.raises(m.MexError, match="same size"):
        m.absoluteOrientationQuaternion(np.zeros((3, 5)), np.zeros((3, 6)))
    with pytest.raises(m.MexError, match="dimension 3"):
        m.absoluteOrientationQuaternion(np.zeros((2, 5)), np.zeros((2, 5)))


def test_R2q_host_helper(pre3, orc):
    rng = np.random.default_rng(1)
    for _ in range(5):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        np.testing.assert_array_equal(pre3.R2q(R), orc.R2q(R))

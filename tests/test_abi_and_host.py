"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/b200msm.h
declares, argument errors surface without a GPU, and the product path has no CPU fallback."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def M():
    import __graft_entry__ as g
    m = g.load_package()
    if not os.path.exists(m.LIB_PATH):
        g.build()
    return m


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "b200msm.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", hdr)
    return sorted({n for n in names if n.startswith(("b200msm_", "mult_pippenger", "MSM"))})


def test_library_exports_every_declared_symbol(M):
    L = ctypes.CDLL(M.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 24
    for name in decl:
        assert hasattr(L, name), f"{name} declared in include/b200msm.h but not exported"
    assert sorted(M.EXPORTED_SYMBOLS) == decl


def test_no_cpu_fallback(M):
    """Without a CUDA device every entry point fails loudly (-2 / RustError / NULL); nothing computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(M.MsmError):
        M.Engine()
    pts = np.zeros((4, 104), np.uint8)
    with pytest.raises(M.MsmError):
        M.multi_scalar_mult_init(pts)
    with pytest.raises(M.MsmError):
        M.YrridContext()
    sc = np.zeros((4, 32), np.uint8)
    with pytest.raises(M.MsmError, match="no CPU path"):
        M.VariableBaseMSM.multi_scalar_mul(pts, sc)
    with pytest.raises(M.MsmError, match="no CPU path"):
        M.VariableBaseMSM.multi_scalar_mul(pts[:0], sc[:0])          # not even the empty MSM is answered on the CPU
    with pytest.raises(M.MsmError):
        M.VariableBase.msm(pts, sc)
    with pytest.raises(M.MsmError):
        M.multi_scalar_mult_blst(pts[:, :96], sc)


def test_bad_arguments(M):
    L = M.lib()
    h = ctypes.c_void_p()
    assert L.b200msm_create(ctypes.byref(h), 7, 0, 0, 0) == -1
    assert L.b200msm_create(None, 0, 0, 0, 0) == -1
    assert L.b200msm_destroy(None) == -1
    assert L.MSMFreeContext(None) == -1
    assert L.MSMRun(None, None, None, 0) == -1
    assert L.MSMAllocContext(0, 1) is None
    out = np.zeros(144, np.uint8)
    assert L.b200msm_variable_base_msm(2, 0, None, 0, 104, None, 0, out.ctypes.data_as(ctypes.c_void_p)) == -1     # curve
    assert L.b200msm_variable_base_msm(0, 0, None, 0, 100, None, 0, out.ctypes.data_as(ctypes.c_void_p)) == -1     # stride
    assert L.b200msm_variable_base_msm(0, 0, None, 0, 104, None, 0, None) == -1                                    # out
    assert L.b200msm_variable_base_msm(0, 0, None, 4, 104, None, 4, out.ctypes.data_as(ctypes.c_void_p)) == -1     # inputs
    assert L.b200msm_variable_base_release() == 0
    with pytest.raises(M.MsmError, match="length mismatch"):
        M.multi_scalar_mult_blst(np.zeros((4, 96), np.uint8), np.zeros((5, 32), np.uint8))
    with pytest.raises(M.MsmError):
        M.multi_scalar_mult(M.MultiScalarMultContext(), np.zeros((4, 104), np.uint8), np.zeros((5, 32), np.uint8))   # length mismatch


def build_c_harness():
    """g++ build of tests/c_harness/msm_harness.cpp (the C++ twin of the reference's tests/msm.rs) against the two .so files."""
    import subprocess
    from oracle import oracle as O
    O.build()
    d = os.path.join(ROOT, "tests", "c_harness")
    exe = os.path.join(d, "msm_harness")
    subprocess.check_call(["g++", "-std=c++17", "-O2", os.path.join(d, "msm_harness.cpp"), "-I" + os.path.join(ROOT, "include"),
                           "-L" + os.path.join(ROOT, "2022-entries_b200"), "-lb200msm", "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                           "-Wl,-rpath," + os.path.join(ROOT, "2022-entries_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-o", exe])
    return exe


def test_c_harness_links_and_fails_loudly_without_gpu(M):
    """The compiled-code client of the C ABI links against every symbol it uses; with no CUDA device it must stop with
    the library's error, not compute anything."""
    import subprocess
    import torch
    exe = build_c_harness()
    if torch.cuda.is_available():
        pytest.skip("CUDA device present: the GPU suite runs the harness for real")
    p = subprocess.run([exe, "8"], capture_output=True, text=True)
    assert p.returncode == 2 and "no CPU path" in p.stderr


def test_window_plan_invariants_and_headline_choice(M):
    """choose_plan (csrc/msm.cu) through b200msm_plan_info, no device: invariants over sizes 2^0..2^30, both curves, memory
    budgets and forced parameters; and the plans DESIGN.md quotes."""
    bits = {0: 253, 1: 255}
    for curve in (0, 1):
        for lg in list(range(0, 31, 2)) + [23, 25, 27]:
            n = 1 << lg
            if n >= 1 << 31:
                continue
            for free in (8 << 30, 40 << 30, 170 << 30):
                for wb, tb, ch in ((0, 0, 1), (0, 1, 1), (0, 1, 4), (0, 3, 1), (7, 0, 1), (16, 1, 2), (23, 0, 1)):
                    p = M.plan_info(n, curve, free, wb, tb, ch)
                    c, W, T, F = p["window_bits"], p["windows"], p["tables"], p["folded_windows"]
                    assert 2 <= c <= 26 and (wb == 0 or c == wb)
                    assert W == -(-bits[curve] // c) and W * c >= bits[curve]
                    assert 1 <= T <= W and T * F >= W and (T - 1) * F < W          # every window has a table, no empty table
                    assert tb == 0 or T <= tb
                    assert p["buckets"] == F << (c - 1) and p["buckets"] <= 1 << 27
                    assert T == 1 or T * n < 1 << 31                               # 31-bit point references
                    assert T == 1 or p["table_bytes"] <= 0.72 * free               # tables fit the memory budget
                    assert p["refs"] == W * n
    # the headline plan: 2^26 points on one B200 fold all 11 windows into one through 11 tables
    p = M.plan_info(1 << 26, 0, 170 << 30)
    assert (p["window_bits"], p["windows"], p["tables"], p["folded_windows"], p["buckets"]) == (23, 11, 11, 1, 1 << 22)
    assert p["table_bytes"] == 11 * (1 << 26) * 144
    # sharded over 8 GPUs the plan is re-derived for n / 8: (23, 11, 1) is the measured-best plan at 2^23 points (35.9 ms against 41.9 /
    # 37.9 ms for c = 22 / 21, profiles/r02_plan_calibration.txt); the one-call form is one table with the chunked reduction charged
    p = M.plan_info(1 << 23, 0, 170 << 30)
    assert (p["window_bits"], p["tables"], p["folded_windows"]) == (23, 11, 1)
    p = M.plan_info(1 << 26, 0, 170 << 30, 0, 1, 4)
    assert (p["window_bits"], p["windows"], p["tables"], p["folded_windows"]) == (20, 13, 1, 13)
    p = M.plan_info(1 << 24, 1, 170 << 30)
    assert (p["window_bits"], p["windows"], p["tables"]) == (22, 12, 12)           # BLS12-381 at 2^24 (profiles/r01_sweep.txt)
    # 2^28: 38.7 GB per table -> 3 tables, 4 folded windows
    p = M.plan_info(1 << 28, 0, 170 << 30)
    assert (p["window_bits"], p["tables"], p["folded_windows"]) == (23, 3, 4)
    with pytest.raises(M.MsmError):
        M.plan_info(0)


def test_window_plan_rejects_combinations_beyond_the_bucket_limit(M):
    """A forced window of 25 / 26 bits with too few tables needs more than 2^27 buckets: an error, not a division by zero
    (choose_plan returned c = 0 and the process died with SIGFPE before)."""
    for wb in (25, 26):
        with pytest.raises(M.MsmError):
            M.plan_info(1 << 20, 0, 170 << 30, wb, 1)
    p = M.plan_info(1 << 20, 0, 170 << 30, 26, 0)           # all tables: one folded window of 2^25 buckets
    assert p["buckets"] == 1 << 25 and p["folded_windows"] == 1
    assert "b200msm_set_bucket_arithmetic" in M.EXPORTED_SYMBOLS

"""CPU checks of the device field / curve headers (csrc/fp.cuh, csrc/ec.cuh) compiled for the host.

The headers are plain C++ apart from the __device__ markers, so g++ builds them (tests/host_emul/emul.cpp) and the
limb-level algorithms (14 x 28-bit reduced radix, R = 2^392, lazy reduction bounds) are compared with the CPU oracle
without a GPU.  The GPU parity tests (test_gpu_parity.py) repeat the comparison on the real kernels.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import ref_bigint as RB

HERE = os.path.join(ROOT, "tests", "host_emul")


def _build_emul(name, flags):
    so = os.path.join(HERE, name)
    src = os.path.join(HERE, "emul.cpp")
    hdrs = [os.path.join(ROOT, "2022-entries_b200", "csrc", h) for h in ("fp.cuh", "fp29.cuh", "ec.cuh", "te.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-x", "c++"] + flags + [src, "-o", so])
    return so


def _hot_flags():
    """The -D switches the hot translation unit of the library is compiled with (csrc/Makefile HOT_DEFS): the host build
    must emulate the same Te::madd."""
    mk = open(os.path.join(ROOT, "2022-entries_b200", "csrc", "Makefile")).read()
    for line in mk.splitlines():
        if line.startswith("HOT_DEFS"):
            return line.split("=", 1)[1].split()
    return []


@pytest.fixture(scope="module", params=["hot", "plain"])
def emul(request):
    """Two host builds: with the hot kernels' switches (what k_accumulate runs) and without (what the cold kernels run)."""
    so = _build_emul("emul.so", _hot_flags()) if request.param == "hot" else _build_emul("emul_plain.so", [])
    L = ctypes.CDLL(so)
    L.emul_fp_mul_raw.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    L.emul_fp29_mul_raw.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    L.emul_te29_op.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    L.emul_fp29_constants.argtypes = [ctypes.c_void_p]
    vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    L.emul_fp_op.argtypes = [i32, i32, vp, vp, vp, sz]
    L.emul_ec_op.argtypes = [i32, i32, vp, vp, sz, i32]
    L.emul_fp_mul_lazy.argtypes = [i32, vp, vp, vp, sz, i32, i32]
    L.emul_constants.argtypes = [i32, vp]
    L.emul_te_op.argtypes = [i32, vp, vp, vp, sz, i32]
    L.emul_te_constants.argtypes = [vp]
    return L


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _limbs(v, n=14, bits=28):
    return [(v >> (bits * i)) & ((1 << bits) - 1) if i < n - 1 else v >> (bits * i) for i in range(n)]


@pytest.mark.parametrize("curve", [0, 1])
def test_constant_tables(emul, curve):
    """Every table in fp.cuh re-derived from p with Python integers."""
    buf = np.zeros(64 + 16 * 14, np.uint32)
    emul.emul_constants(curve, _p(buf))
    p = RB.CURVES[curve]["p"]
    val = lambda a: sum(int(x) << (28 * i) for i, x in enumerate(a))
    assert val(buf[0:14]) == p
    assert val(buf[14:28]) == (1 << 392) % p
    assert val(buf[28:42]) == (1 << 400) % p
    assert val(buf[42:56]) == (1 << 384) % p
    assert int(buf[56]) == (-pow(p, -1, 1 << 28)) % (1 << 28)
    assert int(buf[57]) == pow(p, -1, 1 << 28)
    for k in range(16):
        row = buf[64 + 14 * k: 64 + 14 * (k + 1)]
        assert val(row) == (k + 1) * p                     # limb-dominant form is still exactly K p
        assert all(int(x) >= (1 << 28) - 1 for x in row[:13])
        assert all(int(x) < (1 << 32) for x in row)


def _rand_fp(rng, n, p):
    vals = [int.from_bytes(rng.bytes(48), "little") % p for _ in range(n)]
    return vals


def _pack(vals):
    return np.frombuffer(b"".join(v.to_bytes(48, "little") for v in vals), np.uint8).copy()


@pytest.mark.parametrize("curve", [0, 1])
def test_field_ops(emul, oracle, curve):
    rng = np.random.default_rng(17 + curve)
    p = RB.CURVES[curve]["p"]
    n = 2048
    va, vb = _rand_fp(rng, n, p), _rand_fp(rng, n, p)
    va[:4] = [0, p - 1, 1, p - 1]
    vb[:4] = [5, p - 1, p - 1, 0]
    a, b = _pack(va), _pack(vb)
    r = np.zeros_like(a)
    for op, oop in ((0, oracle.FP_MUL), (1, oracle.FP_ADD), (2, oracle.FP_SUB), (3, oracle.FP_SQR)):
        emul.emul_fp_op(curve, op, _p(a), _p(b), _p(r), n)
        assert np.array_equal(r.reshape(-1, 48), oracle.fp_op(oop, a, b, curve)), (curve, op)
    emul.emul_fp_op(curve, 4, _p(a[48:48 * 9]), _p(a[48:48 * 9]), _p(r), 8)
    assert np.array_equal(r[:48 * 8].reshape(-1, 48), oracle.fp_op(oracle.FP_INV, a[48:48 * 9], None, curve))
    emul.emul_fp_op(curve, 5, _p(a), _p(b), _p(r), n)
    assert np.array_equal(r, a)


@pytest.mark.parametrize("curve,k1,k2", [(0, 13, 13), (0, 10, 2), (1, 13, 13), (1, 8, 13)])
def test_lazy_multiply_bound(emul, oracle, curve, k1, k2):
    """(a + k1 p)(b + k2 p) stays below 2p after the Montgomery multiply: the bounds ec.cuh relies on."""
    rng = np.random.default_rng(5)
    p = RB.CURVES[curve]["p"]
    n = 1024
    va, vb = _rand_fp(rng, n, p), _rand_fp(rng, n, p)
    va[0], vb[0] = p - 1, p - 1
    a, b = _pack(va), _pack(vb)
    r = np.zeros_like(a)
    assert emul.emul_fp_mul_lazy(curve, _p(a), _p(b), _p(r), n, k1, k2) == 0
    assert np.array_equal(r.reshape(-1, 48), oracle.fp_op(oracle.FP_MUL, a, b, curve))


@pytest.mark.parametrize("la,lb", [(3, 1), (3, 2), (2, 3), (1, 1), (6, 1)])
def test_multiplier_accepts_unnormalised_limbs(emul, la, lb):
    """Te::madd (TE_LAZY_NORM) hands the multiplier sums whose limbs were not carried: limbs 0..12 up to la * 2^28 and
    lb * 2^28 with la * lb <= 6 (te.cuh).  Worst case (every limb at its bound) and random fills: the result must be the
    Montgomery product of the VALUES, normalised -- i.e. no 64-bit column overflowed."""
    rng = np.random.default_rng(la * 10 + lb)
    p = RB.CURVES[0]["p"]
    n = 64
    A = rng.integers(0, la << 28, size=(n, 14), dtype=np.uint64).astype(np.uint32)
    B = rng.integers(0, lb << 28, size=(n, 14), dtype=np.uint64).astype(np.uint32)
    A[0, :13] = (la << 28) - 1
    B[0, :13] = (lb << 28) - 1
    A[1, :13] = (la << 28) - 1
    B[2, :13] = (lb << 28) - 1
    A[:, 13] = rng.integers(0, 1 << 17, size=n)            # the top limb is bounded by the value, not by the carry state
    B[:, 13] = rng.integers(0, 1 << 17, size=n)
    R = np.zeros((n, 14), np.uint32)
    emul.emul_fp_mul_raw(0, _p(A), _p(B), _p(R), n)
    rinv = pow(1 << 392, -1, p)
    for i in range(n):
        va = sum(int(A[i, j]) << (28 * j) for j in range(14))
        vb = sum(int(B[i, j]) << (28 * j) for j in range(14))
        vr = sum(int(R[i, j]) << (28 * j) for j in range(14))
        assert all(int(R[i, j]) < (1 << 28) for j in range(13)), "result not normalised"
        assert vr % p == va * vb * rinv % p
        assert vr < va * vb // (1 << 392) + p + 1


def test_fp29_constants(emul):
    """csrc/fp29.cuh tables (13 x 29-bit limbs, R = 2^406) and the conversion constants of fp.cuh, re-derived from p."""
    p = RB.CURVES[0]["p"]
    out = np.zeros(192, np.uint32)
    emul.emul_fp29_constants(_p(out))
    val = lambda off, n=13, bits=29: sum(int(out[off + i]) << (bits * i) for i in range(n))
    assert val(0) == p and all(int(out[i]) < (1 << 29) for i in range(13))
    assert val(13) == pow(2, 406, p) and val(26) == pow(2, 392, p)
    assert (-pow(p, -1, 1 << 29)) % (1 << 29) == (1 << 29) - 1 and p % (1 << 29) == 1
    for k in range(8):                                   # K p in limb-dominant form: same value, every low limb >= 2^29 - 1
        assert val(39 + 13 * k) == (k + 1) * p
        assert all(int(out[39 + 13 * k + i]) >= (2 << 29) - 2 for i in range(12))
    assert val(160, 14, 28) == pow(2, 406, p) and val(174, 14, 28) == pow(2, 378, p)


def test_fp29_multiplier(emul):
    """Fp29::mul = a b / 2^406 mod p on normalised operands up to the largest representatives the mixed add forms (7p x 6p:
    the top limb takes the excess), worst-case limbs included: no 64-bit column overflows, result normalised and < 1.1 p."""
    rng = np.random.default_rng(29)
    p = RB.CURVES[0]["p"]
    n = 256
    def rand_rep(k):                                     # a representative of a random residue below k p, as 13 x 29 limbs
        v = int.from_bytes(rng.bytes(48), "little") % p + int(rng.integers(0, k)) * p
        return v
    A = np.zeros((n, 13), np.uint32); B = np.zeros((n, 13), np.uint32)
    va, vb = [], []
    for i in range(n):
        a, b = rand_rep(7), rand_rep(6)
        if i == 0:
            a, b = 7 * p - 1, 6 * p - 1
        if i == 1:                                       # every low limb at its maximum ...
            a = ((7 * p) >> 348 << 348) - 1
            b = ((6 * p) >> 348 << 348) - 1
        over = 8 if i in (1, 2) else 0                   # ... and beyond: the parallel carry leaves limbs up to 2^29 + 7
        for j in range(13):
            A[i, j] = ((a >> (29 * j)) & ((1 << 29) - 1)) + (over if j else 0) if j < 12 else a >> 348
            B[i, j] = ((b >> (29 * j)) & ((1 << 29) - 1)) + (over if j else 0) if j < 12 else b >> 348
        va.append(sum(int(A[i, j]) << (29 * j) for j in range(13))); vb.append(sum(int(B[i, j]) << (29 * j) for j in range(13)))
    R = np.zeros((n, 13), np.uint32)
    emul.emul_fp29_mul_raw(_p(A), _p(B), _p(R), n)
    rinv = pow(1 << 406, -1, p)
    for i in range(n):
        vr = sum(int(R[i, j]) << (29 * j) for j in range(13))
        assert all(int(R[i, j]) < (1 << 29) for j in range(12))
        assert vr % p == va[i] * vb[i] * rinv % p
        assert vr < p + p // 8


def test_te29_group_law(emul, oracle):
    """Te29::madd (the accumulate kernel's mixed add, 13 x 29-bit field, entries in 2^406 form) == the oracle's Jacobian
    arithmetic: P + Q chains, P + P, P - P through the neutral element, negated entries; representation round trip."""
    O = oracle
    n = 48
    A, B = O.gen_points(n, seed=15), O.gen_points(n, seed=19)
    B[0] = A[0]
    B[1] = O.aff_neg(A[1:2])[0]
    a96, b96 = np.ascontiguousarray(A[:, :96]), np.ascontiguousarray(B[:, :96])
    out = np.zeros((n, 104), np.uint8)
    jacA = _jac_from_aff(O, A, 0)
    for op, reps in ((0, 1), (0, 7), (3, 1), (3, 5), (0, 0)):
        emul.emul_te29_op(op, _p(a96), _p(b96), _p(out), n, reps)
        j = jacA.copy()
        for _ in range(reps):
            j = O.jac_add_mixed(j, B if op == 0 else O.aff_neg(B))
        want = O.jac_to_affine(j)
        assert np.array_equal(out[:, 96], want[:, 96]), (op, reps)
        fin = want[:, 96] == 0
        assert np.array_equal(out[fin, :96], want[fin, :96]), (op, reps)


def _xyzz_from_aff(O, aff, curve):
    one = np.frombuffer(O.constants(curve)["R"].to_bytes(48, "little"), np.uint8)
    out = np.zeros((len(aff), 192), np.uint8)
    out[:, :96] = aff[:, :96]
    out[:, 96:144] = one
    out[:, 144:] = one
    return out


def _jac_from_aff(O, aff, curve):
    one = np.frombuffer(O.constants(curve)["R"].to_bytes(48, "little"), np.uint8)
    out = np.zeros((len(aff), 144), np.uint8)
    out[:, :96] = aff[:, :96]
    out[:, 96:] = one
    return out


@pytest.mark.parametrize("curve", [0, 1])
def test_curve_ops(emul, oracle, curve):
    O = oracle
    n = 64
    A, B = O.gen_points(n, seed=5, curve=curve), O.gen_points(n, seed=9, curve=curve)
    B[0] = A[0]                                   # P + P -> doubling branch
    B[1] = O.aff_neg(A[1:2], curve)[0]            # P - P -> infinity
    acc = _xyzz_from_aff(O, A, curve)
    acc[2] = 0                                    # infinity + P
    jac = _jac_from_aff(O, A, curve)
    jac[2, 96:] = 0
    b96 = np.ascontiguousarray(B[:, :96])
    got = acc.copy()
    emul.emul_ec_op(curve, 0, _p(got), _p(b96), n, 1)
    assert np.array_equal(O.xyzz_to_affine(got, curve), O.jac_to_affine(O.jac_add_mixed(jac, B, curve), curve))
    got9 = _xyzz_from_aff(O, A, curve)
    emul.emul_ec_op(curve, 0, _p(got9), _p(b96), n, 9)
    j = _jac_from_aff(O, A, curve)
    for _ in range(9):
        j = O.jac_add_mixed(j, B, curve)
    assert np.array_equal(O.xyzz_to_affine(got9, curve), O.jac_to_affine(j, curve))
    lhs, other = got9.copy(), got9.copy()
    lhs[3] = 0
    other[4] = 0
    jl, jo = j.copy(), j.copy()
    jl[3, 96:] = 0
    jo[4, 96:] = 0
    emul.emul_ec_op(curve, 1, _p(lhs), _p(other), n, 1)
    assert np.array_equal(O.xyzz_to_affine(lhs, curve), O.jac_to_affine(O.jac_add(jl, jo, curve), curve))
    d = got9.copy()
    emul.emul_ec_op(curve, 2, _p(d), None, n, 23)
    jd = j.copy()
    for _ in range(23):
        jd = O.jac_double(jd, curve)
    assert np.array_equal(O.xyzz_to_affine(d, curve), O.jac_to_affine(jd, curve))
    da = np.zeros((n, 192), np.uint8)
    a96 = np.ascontiguousarray(A[:, :96])
    emul.emul_ec_op(curve, 3, _p(da), _p(a96), n, 1)
    assert np.array_equal(O.xyzz_to_affine(da, curve), O.jac_to_affine(O.jac_double(_jac_from_aff(O, A, curve), curve), curve))


def test_curve_ops_reproduce_rfc9380_cofactor_clearing(emul, oracle, bls381_hash_to_curve):
    """The XYZZ arithmetic of csrc/ec.cuh (host build) on the reference's own BLS12-381 known answers: arkworks' RFC 9380
    hash-to-curve vectors give Q0, Q1 (on the curve, outside G1, full-size coordinates) and P = h_eff * Q0 + h_eff * Q1.
    h_eff * Q by double-and-add with dbl / madd, the two halves joined with the full XYZZ add."""
    O = oracle
    vectors, h = bls381_hash_to_curve
    Q = np.concatenate([b for b, _ in vectors])                      # Q0, Q1 of vector 0, of vector 1, ...
    n = len(Q)
    q96 = np.ascontiguousarray(Q[:, :96])
    acc = _xyzz_from_aff(O, Q, 1)                                    # top bit of h_eff
    for bit in bin(h)[3:]:
        emul.emul_ec_op(1, 2, _p(acc), None, n, 1)
        if bit == "1":
            emul.emul_ec_op(1, 0, _p(acc), _p(q96), n, 1)
    lhs, rhs = np.ascontiguousarray(acc[0::2]), np.ascontiguousarray(acc[1::2])
    emul.emul_ec_op(1, 1, _p(lhs), _p(rhs), n // 2, 1)
    want = np.stack([w for _, w in vectors])
    assert np.array_equal(O.xyzz_to_affine(lhs, 1), want)


def test_te_constants(emul):
    """Twisted-Edwards parameters of BLS12-377 G1 re-derived from p: s = 1/sqrt 3, A = -3s, B = s, a1 = (A+2)/B,
    d1 = (A-2)/B, f = sqrt(-a1), d = -d1/a1 (= arkworks' COEFF_D, P1B/nickray/driver/bls12-377/src/curves/g1.rs:93)."""
    buf = np.zeros(70, np.uint32)
    emul.emul_te_constants(_p(buf))
    p = RB.CURVES[0]["p"]
    R = (1 << 392) % p
    val = lambda a: sum(int(x) << (28 * i) for i, x in enumerate(a)) * pow(R, -1, p) % p
    k, f, s, s3, cy = (val(buf[14 * i: 14 * i + 14]) for i in range(5))
    assert 3 * s * s % p == 1 and s * s3 % p == 1
    A, B = (-3 * s) % p, s
    a1, d1 = (A + 2) * pow(B, -1, p) % p, (A - 2) * pow(B, -1, p) % p
    d = (-d1 * pow(a1, -1, p)) % p
    assert d == 122268283598675559488486339158635529096981886914877139579534153582033676785385790730042363341236035746924960903179
    assert k == 2 * d % p and f * f % p == (-a1) % p and cy == f * s3 % p


def test_te_group_law(emul, oracle):
    """Mixed add, full add, doubling and subtraction in twisted-Edwards coordinates == the oracle's Jacobian arithmetic,
    including P + P, P - P and chains that pass through the neutral element."""
    O = oracle
    n = 48
    A, B = O.gen_points(n, seed=5), O.gen_points(n, seed=9)
    B[0] = A[0]                                   # P + P through the unified formula
    B[1] = O.aff_neg(A[1:2])[0]                   # P - P -> neutral, then neutral + B ...
    a96, b96 = np.ascontiguousarray(A[:, :96]), np.ascontiguousarray(B[:, :96])
    out = np.zeros((n, 104), np.uint8)
    jacA, jacB = _jac_from_aff(O, A, 0), _jac_from_aff(O, B, 0)
    for op, reps in ((0, 1), (0, 5), (1, 3), (2, 23), (3, 1), (3, 4), (0, 0)):
        emul.emul_te_op(op, _p(a96), _p(b96), _p(out), n, reps)
        j = jacA.copy()
        for _ in range(reps):
            if op in (0, 1):
                j = O.jac_add_mixed(j, B)
            elif op == 2:
                j = O.jac_double(j)
            else:
                j = O.jac_add_mixed(j, O.aff_neg(B))
        want = O.jac_to_affine(j)
        assert np.array_equal(out[:, 96], want[:, 96]), (op, reps)
        fin = want[:, 96] == 0
        assert np.array_equal(out[fin, :96], want[fin, :96]), (op, reps)


# ---------------------------------------------------------------------------------------------------------------
# FP64-pipe field (csrc/fpd.cuh) and accumulator (te.cuh TeD), host build with the FPU rounding toward zero
@pytest.fixture(scope="module")
def emuld():
    so = os.path.join(HERE, "emul_d.so")
    src = os.path.join(HERE, "emul_d.cpp")
    hdrs = [os.path.join(ROOT, "2022-entries_b200", "csrc", h) for h in ("fp.cuh", "fpd.cuh", "ec.cuh", "te.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-frounding-math", "-fPIC", "-shared", "-x", "c++", src, "-o", so])
    L = ctypes.CDLL(so)
    vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    L.emuld_fp_op.argtypes = [i32, vp, vp, vp, sz]
    L.emuld_fp_mul_lazy.argtypes = [vp, vp, vp, sz, i32, i32]
    L.emuld_is_zero.argtypes = [vp, i32, i32]
    L.emuld_te_op.argtypes = [i32, vp, vp, vp, sz, i32]
    L.emuld_constants.argtypes = [vp]
    return L


def test_fp64_field_constants(emuld):
    c = np.zeros(24 + 128, np.uint64)
    emuld.emuld_constants(_p(c))
    p = RB.CURVES[0]["p"]
    val = lambda a: sum(int(x) << (48 * i) for i, x in enumerate(a))
    assert val(c[0:8]) == p and val(c[8:16]) == (1 << 384) % p
    assert int(c[16]) == (-pow(p, -1, 1 << 48)) % (1 << 48) and int(c[17]) == pow(p, -1, 1 << 48)
    for k in range(16):
        row = c[24 + 8 * k: 32 + 8 * k]
        assert val(row) == (k + 1) * p and all(int(x) >= (1 << 48) - 1 for x in row[:7])


def test_fp64_field_ops(emuld, oracle):
    rng = np.random.default_rng(23)
    p = RB.CURVES[0]["p"]
    n = 2048
    va, vb = _rand_fp(rng, n, p), _rand_fp(rng, n, p)
    va[:4] = [0, p - 1, 1, p - 1]
    vb[:4] = [5, p - 1, p - 1, 0]
    a, b = _pack(va), _pack(vb)
    r = np.zeros_like(a)
    for op, oop in ((0, oracle.FP_MUL), (1, oracle.FP_ADD), (2, oracle.FP_SUB), (3, oracle.FP_SQR)):
        emuld.emuld_fp_op(op, _p(a), _p(b), _p(r), n)
        assert np.array_equal(r.reshape(-1, 48), oracle.fp_op(oop, a, b, 0)), op
    emuld.emuld_fp_op(4, _p(a[48:48 * 9]), _p(a[48:48 * 9]), _p(r), 8)
    assert np.array_equal(r[:48 * 8].reshape(-1, 48), oracle.fp_op(oracle.FP_INV, a[48:48 * 9], None, 0))
    emuld.emuld_fp_op(5, _p(a), _p(b), _p(r), n)
    assert np.array_equal(r, a)
    # lazy operands: (a + k1 p)(b + k2 p) with limbs up to 4 units stays normalised and below 2p (ka kb <= 152)
    for k1, k2 in ((12, 12), (7, 6), (9, 2), (0, 0)):
        assert emuld.emuld_fp_mul_lazy(_p(a), _p(b), _p(r), n, k1, k2) == 0
        assert np.array_equal(r.reshape(-1, 48), oracle.fp_op(oracle.FP_MUL, a, b, 0))
    z, nz = _pack([0]), _pack([12345])
    for k in (0, 1, 5, 12):
        assert emuld.emuld_is_zero(_p(z), k, 0) == 1 and emuld.emuld_is_zero(_p(z), k, 1) == 1 and emuld.emuld_is_zero(_p(nz), k, 0) == 0


def test_fp64_te_accumulator(emuld, oracle):
    """The accumulate kernel's inner loop arithmetic (TeD::madd chains from the neutral element, negated entries,
    P + P, P - P) == the oracle."""
    O = oracle
    n = 48
    A, B = O.gen_points(n, seed=25), O.gen_points(n, seed=29)
    B[0] = A[0]
    B[1] = O.aff_neg(A[1:2])[0]
    a96, b96 = np.ascontiguousarray(A[:, :96]), np.ascontiguousarray(B[:, :96])
    out = np.zeros((n, 104), np.uint8)
    jacA = _jac_from_aff(O, A, 0)
    for op, reps in ((0, 1), (0, 9), (3, 1), (3, 6), (0, 0)):
        emuld.emuld_te_op(op, _p(a96), _p(b96), _p(out), n, reps)
        j = jacA.copy()
        for _ in range(reps):
            j = O.jac_add_mixed(j, B if op == 0 else O.aff_neg(B))
        want = O.jac_to_affine(j)
        assert np.array_equal(out[:, 96], want[:, 96]), (op, reps)
        fin = want[:, 96] == 0
        assert np.array_equal(out[fin, :96], want[fin, :96]), (op, reps)


def test_copy_pool_parallel_range_copy():
    """csrc/copy_pool.h (the host threads behind the staged upload of pageable arrays): odd sizes and offsets, 1-16 threads."""
    exe = os.path.join(HERE, "pool_test")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", os.path.join(HERE, "pool_test.cpp"), "-o", exe])
    for threads in (1, 3, 8, 16):
        p = subprocess.run([exe, str(threads)], capture_output=True, text=True, timeout=120)
        assert p.returncode == 0 and p.stdout.strip() == "ok", (threads, p.stdout, p.stderr)


# ---------------------------------------------------------------------------------------------------------------
# a4 on the host: csrc/digits.cuh (the arithmetic of k_digits / k_digits_bin) against Python integers
@pytest.fixture(scope="module")
def digits():
    so = os.path.join(HERE, "digits_emul.so")
    src = os.path.join(HERE, "digits_emul.cpp")
    hdrs = [os.path.join(ROOT, "2022-entries_b200", "csrc", h) for h in ("digits.cuh", "fp.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-x", "c++", src, "-o", so])
    L = ctypes.CDLL(so)
    vp, sz, i32, u32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32
    L.emul_digits.argtypes = [i32, vp, sz, u32, u32, i32, i32, i32, i32, vp, vp, vp, vp]
    return L


@pytest.mark.parametrize("curve", [0, 1])
def test_signed_digit_decomposition_recomposes(digits, curve):
    """For every window width c = 2..26 and table counts T in {1, 3, W}: the (key, reference) pairs of a scalar name each
    window at most once, stay inside the bucket and table ranges, and  sum +-d_w 2^(c w) = s (mod r)  with the sign carried by
    bit 31 of the reference -- for random scalars, scalars at and above r (reduced first) and the top-bit fold."""
    from oracle import oracle as O
    r = O.constants(curve)["r"]
    bits = 253 if curve == 0 else 255
    rng = np.random.default_rng(5 + curve)
    vals = [0, 1, 2, r - 1, r, r + 1, (r - 1) // 2, (r + 1) // 2, (1 << (bits - 1)) - 1, 1 << (bits - 1), (1 << (bits - 1)) + 1,
            (1 << bits) - 1, (1 << 256) - 1, 2 * r - 1, 2 * r, 2 * r + 5]
    if curve == 0:
        vals += [13 * r + 7, 8 * r, 4 * r - 1]                      # 2^256 < 14 r on BLS12-377
    vals += [int.from_bytes(rng.bytes(32), "little") % r for _ in range(200)]
    vals += [int.from_bytes(rng.bytes(32), "little") for _ in range(40)]        # unreduced 256-bit values
    vals = [v for v in vals if v < (1 << 256)]
    n = len(vals)
    sc = np.zeros((n, 8), np.uint32)
    for i, v in enumerate(vals):
        sc[i] = np.frombuffer(int(v).to_bytes(32, "little"), np.uint32)
    first, stride = 1000, 1 << 20
    keys, refs, wins = np.zeros((n, 32), np.uint32), np.zeros((n, 32), np.uint32), np.zeros((n, 32), np.int32)
    counts = np.zeros(n, np.int32)
    for c in range(2, 27):
        W = (bits + c - 1) // c
        for T in sorted({1, min(3, W), W}):
            F = (W + T - 1) // T
            half = 1 << (c - 1)
            for maxw16 in ((0, 1) if W <= 16 else (0,)):
                if W > 32:
                    continue                                         # the emulation buffers hold 32 pairs per scalar
                digits.emul_digits(curve, _p(sc), n, first, stride, c, W, F, maxw16, _p(keys), _p(refs), _p(wins), _p(counts))
                for i, v in enumerate(vals):
                    total, seen = 0, set()
                    for k in range(counts[i]):
                        key, ref = int(keys[i, k]), int(refs[i, k])
                        f, d = divmod(key, half)
                        d += 1
                        neg, idx = ref >> 31, ref & 0x7FFFFFFF
                        j, gi = divmod(idx, stride)
                        w = j * F + f
                        assert gi == first + i and f < F and j < T and w < W and 1 <= d <= half and w not in seen, (c, T, v)
                        if maxw16:
                            assert int(wins[i, k]) == w
                        seen.add(w)
                        total += (-d if neg else d) << (c * w)
                    assert total % r == v % r, (curve, c, T, hex(v))

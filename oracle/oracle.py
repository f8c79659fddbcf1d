"""ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py)."""
import ctypes
import os
import struct
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

AFF_BYTES, JAC_BYTES, FP_BYTES, SCALAR_BYTES = 104, 144, 48, 32


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        vp, sz, i32, u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64
        L.oracle_ark_window_bits.argtypes = [u64]
        L.oracle_ark_msm.argtypes = [i32, vp, vp, sz, vp, i32, i32]
        L.oracle_naive_msm.argtypes = [i32, vp, vp, sz, vp]
        L.oracle_jac_to_affine.argtypes = [i32, vp, sz, vp]
        L.oracle_xyzz_to_affine.argtypes = [i32, vp, sz, vp]
        L.oracle_jac_eq.argtypes = [i32, vp, vp]
        L.oracle_on_curve.argtypes = [i32, vp, sz]
        L.oracle_jac_add_mixed_n.argtypes = [i32, vp, vp, sz]
        L.oracle_jac_add_n.argtypes = [i32, vp, vp, sz]
        L.oracle_jac_double_n.argtypes = [i32, vp, sz]
        L.oracle_aff_neg_n.argtypes = [i32, vp, sz]
        L.oracle_fp_op.argtypes = [i32, i32, vp, vp, vp, sz]
        L.oracle_fp_convert.argtypes = [i32, i32, vp, vp, sz]
        L.oracle_constants.argtypes = [i32, vp]
        L.oracle_decompress.argtypes = [i32, vp, sz, vp]
        L.oracle_compress.argtypes = [i32, vp, sz, vp]
        L.oracle_gen_points.argtypes = [i32, u64, sz, vp]
        L.oracle_gen_scalars.argtypes = [i32, u64, sz, vp]
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _u8(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8).reshape(-1)


def window_bits(n):
    return lib().oracle_ark_window_bits(n)


def ark_msm(bases, scalars, curve=0, threads=None, c=0):
    """arkworks VariableBaseMSM::multi_scalar_mul -> 144-byte Jacobian (np.uint8[144])."""
    b, s = _u8(bases), _u8(scalars)
    n = min(b.size // AFF_BYTES, s.size // SCALAR_BYTES)
    out = np.zeros(JAC_BYTES, np.uint8)
    rc = lib().oracle_ark_msm(curve, _ptr(b), _ptr(s), n, _ptr(out), threads or os.cpu_count() or 1, c)
    assert rc == 0
    return out


def naive_msm(bases, scalars, curve=0):
    b, s = _u8(bases), _u8(scalars)
    n = min(b.size // AFF_BYTES, s.size // SCALAR_BYTES)
    out = np.zeros(JAC_BYTES, np.uint8)
    assert lib().oracle_naive_msm(curve, _ptr(b), _ptr(s), n, _ptr(out)) == 0
    return out


def jac_to_affine(jac, curve=0):
    j = _u8(jac)
    n = j.size // JAC_BYTES
    out = np.zeros(n * AFF_BYTES, np.uint8)
    assert lib().oracle_jac_to_affine(curve, _ptr(j), n, _ptr(out)) == 0
    return out.reshape(n, AFF_BYTES)


def xyzz_to_affine(xyzz, curve=0):
    q = _u8(xyzz)
    n = q.size // (4 * FP_BYTES)
    out = np.zeros(n * AFF_BYTES, np.uint8)
    assert lib().oracle_xyzz_to_affine(curve, _ptr(q), n, _ptr(out)) == 0
    return out.reshape(n, AFF_BYTES)


def jac_eq(a, b, curve=0):
    a, b = _u8(a), _u8(b)
    return bool(lib().oracle_jac_eq(curve, _ptr(a), _ptr(b)))


def on_curve(aff, curve=0):
    a = _u8(aff)
    return bool(lib().oracle_on_curve(curve, _ptr(a), a.size // AFF_BYTES))


def jac_add_mixed(acc, aff, curve=0):
    acc = _u8(acc).copy()
    a = _u8(aff)
    assert lib().oracle_jac_add_mixed_n(curve, _ptr(acc), _ptr(a), acc.size // JAC_BYTES) == 0
    return acc.reshape(-1, JAC_BYTES)


def jac_add(acc, other, curve=0):
    acc = _u8(acc).copy()
    o = _u8(other)
    assert lib().oracle_jac_add_n(curve, _ptr(acc), _ptr(o), acc.size // JAC_BYTES) == 0
    return acc.reshape(-1, JAC_BYTES)


def jac_double(acc, curve=0):
    acc = _u8(acc).copy()
    assert lib().oracle_jac_double_n(curve, _ptr(acc), acc.size // JAC_BYTES) == 0
    return acc.reshape(-1, JAC_BYTES)


def aff_neg(aff, curve=0):
    a = _u8(aff).copy()
    assert lib().oracle_aff_neg_n(curve, _ptr(a), a.size // AFF_BYTES) == 0
    return a.reshape(-1, AFF_BYTES)


FP_MUL, FP_ADD, FP_SUB, FP_SQR, FP_INV = range(5)


def fp_op(op, a, b=None, curve=0):
    a = _u8(a)
    b = a if b is None else _u8(b)
    out = np.zeros(a.size, np.uint8)
    assert lib().oracle_fp_op(curve, op, _ptr(a), _ptr(b), _ptr(out), a.size // FP_BYTES) == 0
    return out.reshape(-1, FP_BYTES)


def fp_convert(a, to_mont, curve=0):
    a = _u8(a)
    out = np.zeros(a.size, np.uint8)
    assert lib().oracle_fp_convert(curve, int(to_mont), _ptr(a), _ptr(out), a.size // FP_BYTES) == 0
    return out.reshape(-1, FP_BYTES)


def constants(curve=0):
    buf = np.zeros(328, np.uint8)
    assert lib().oracle_constants(curve, _ptr(buf)) == 0
    raw = buf.tobytes()
    f = lambda o, n: int.from_bytes(raw[o:o + n], "little")
    return dict(p=f(0, 48), R=f(48, 48), R2=f(96, 48), pinv=f(144, 8), r=f(152, 32), gx=f(184, 48), gy=f(232, 48),
                b=f(280, 48))


def decompress(buf, curve=0):
    b = _u8(np.frombuffer(bytes(buf), np.uint8))
    n = b.size // 48
    out = np.zeros(n * AFF_BYTES, np.uint8)
    rc = lib().oracle_decompress(curve, _ptr(b), n, _ptr(out))
    assert rc == 0, rc
    return out.reshape(n, AFF_BYTES)


def compress(aff, curve=0):
    a = _u8(aff)
    n = a.size // AFF_BYTES
    out = np.zeros(n * 48, np.uint8)
    assert lib().oracle_compress(curve, _ptr(a), n, _ptr(out)) == 0
    return out.reshape(n, 48)


def gen_points(n, seed=1, curve=0):
    out = np.zeros(n * AFF_BYTES, np.uint8)
    assert lib().oracle_gen_points(curve, seed, n, _ptr(out)) == 0
    return out.reshape(n, AFF_BYTES)


def gen_scalars(n, seed=2, curve=0):
    out = np.zeros(n * SCALAR_BYTES, np.uint8)
    assert lib().oracle_gen_scalars(curve, seed, n, _ptr(out)) == 0
    return out.reshape(n, SCALAR_BYTES)


def read_golden(points_path, scalars_path, result_path=None):
    """Parse the prize-7 files: list of (compressed points bytes, scalars np.uint8[n,32], result bytes|None)."""
    pts, scs = open(points_path, "rb").read(), open(scalars_path, "rb").read()
    res = open(result_path, "rb").read() if result_path else None
    out, po, so, k = [], 0, 0, 0
    while po < len(pts):
        (n,) = struct.unpack_from("<Q", pts, po)
        (m,) = struct.unpack_from("<Q", scs, so)
        assert n == m
        P = pts[po + 8: po + 8 + 48 * n]
        S = np.frombuffer(scs[so + 8: so + 8 + 32 * n], np.uint8).reshape(n, 32)
        out.append((P, S, res[48 * k: 48 * k + 48] if res else None))
        po += 8 + 48 * n
        so += 8 + 32 * n
        k += 1
    return out

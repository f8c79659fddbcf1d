"""GPU parity tests proper (-m gpu): every call goes through the C ABI of libb200msm.so and is compared with the
CPU oracle bit-exactly after normalising to canonical affine (the MSM result is a unique group element)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import __graft_entry__ as g
    return g.load_package()


def xyzz_from_aff(O, aff, curve):
    one = np.frombuffer(O.constants(curve)["R"].to_bytes(48, "little"), np.uint8)
    out = np.zeros((len(aff), 192), np.uint8)
    out[:, :96] = aff[:, :96]
    out[:, 96:144] = one
    out[:, 144:] = one
    out[aff[:, 96] != 0] = 0
    return out


def jac_from_aff(O, aff, curve):
    one = np.frombuffer(O.constants(curve)["R"].to_bytes(48, "little"), np.uint8)
    out = np.zeros((len(aff), 144), np.uint8)
    out[:, :96] = aff[:, :96]
    out[:, 96:] = one
    out[aff[:, 96] != 0, 96:] = 0
    return out


@pytest.mark.parametrize("curve", [0, 1])
def test_field_ops_match_oracle(M, oracle, curve):
    rng = np.random.default_rng(3)
    p = oracle.constants(curve)["p"]
    n = 4096
    vals = [int.from_bytes(rng.bytes(48), "little") % p for _ in range(2 * n)]
    vals[0], vals[1], vals[2], vals[n], vals[n + 1] = 0, p - 1, 1, p - 1, p - 1
    a = np.frombuffer(b"".join(v.to_bytes(48, "little") for v in vals[:n]), np.uint8).copy()
    b = np.frombuffer(b"".join(v.to_bytes(48, "little") for v in vals[n:]), np.uint8).copy()
    for op, oop in ((0, oracle.FP_MUL), (1, oracle.FP_ADD), (2, oracle.FP_SUB), (3, oracle.FP_SQR)):
        assert np.array_equal(M.test_fp_op(curve, op, a, b), oracle.fp_op(oop, a, b, curve)), (curve, op)
    assert np.array_equal(M.test_fp_op(curve, 4, a[48:48 * 33]), oracle.fp_op(oracle.FP_INV, a[48:48 * 33], None, curve))
    assert np.array_equal(M.test_fp_op(curve, 5, a), a.reshape(-1, 48))      # ABI -> device representation -> ABI


@pytest.mark.parametrize("curve", [0, 1])
def test_curve_ops_match_oracle(M, oracle, curve):
    O = oracle
    n = 256
    A, B = O.gen_points(n, seed=5, curve=curve), O.gen_points(n, seed=9, curve=curve)
    B[0] = A[0]                                   # P + P
    B[1] = O.aff_neg(A[1:2], curve)[0]            # P - P
    acc = xyzz_from_aff(O, A, curve)
    acc[2] = 0                                    # infinity + P
    jac = jac_from_aff(O, A, curve)
    jac[2, 96:] = 0
    b96 = np.ascontiguousarray(B[:, :96])
    got = M.test_ec_op(curve, 0, acc, b96)
    assert np.array_equal(O.xyzz_to_affine(got, curve), O.jac_to_affine(O.jac_add_mixed(jac, B, curve), curve))
    # 7 mixed adds in a row: the lazy-reduction bounds must hold across iterations
    got7 = M.test_ec_op(curve, 0, xyzz_from_aff(O, A, curve), b96, reps=7)
    j = jac_from_aff(O, A, curve)
    for _ in range(7):
        j = O.jac_add_mixed(j, B, curve)
    assert np.array_equal(O.xyzz_to_affine(got7, curve), O.jac_to_affine(j, curve))
    # full add, incl. equal operands (doubling) and infinity on either side
    other = got7.copy()
    lhs = got7.copy()
    lhs[3] = 0
    other[4] = 0
    jl, jo = j.copy(), j.copy()
    jl[3, 96:] = 0
    jo[4, 96:] = 0
    got_add = M.test_ec_op(curve, 1, lhs, other)
    assert np.array_equal(O.xyzz_to_affine(got_add, curve), O.jac_to_affine(O.jac_add(jl, jo, curve), curve))
    # 23 doublings (one base-table step) and affine doubling
    got_d = M.test_ec_op(curve, 2, got7, reps=23)
    jd = j.copy()
    for _ in range(23):
        jd = O.jac_double(jd, curve)
    assert np.array_equal(O.xyzz_to_affine(got_d, curve), O.jac_to_affine(jd, curve))
    got_da = M.test_ec_op(curve, 3, np.zeros((n, 192), np.uint8), np.ascontiguousarray(A[:, :96]))
    assert np.array_equal(O.xyzz_to_affine(got_da, curve), O.jac_to_affine(O.jac_double(jac_from_aff(O, A, curve), curve), curve))


@pytest.mark.parametrize("k", range(5))
def test_golden_vectors_through_harness_api(M, oracle, golden_instances, k):
    """The five arkworks-generated instances, via mult_pippenger_init / mult_pippenger_inf."""
    g = golden_instances[k]
    ctx = M.multi_scalar_mult_init(g["aff"])
    out = M.multi_scalar_mult(ctx, g["aff"], g["scalars"])
    ctx.free()
    assert oracle.compress(oracle.jac_to_affine(out[0])).tobytes() == g["result"]


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()

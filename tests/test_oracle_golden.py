"""Pins the CPU oracle against the reference's own known-answer vectors.

The vectors were produced by arkworks VariableBaseMSM::multi_scalar_mul
(open-division/prize7-msm-mobile/gnark/src/lib.rs:245-283) and are checked by the reference in
gnark/msm/bls12-377/zprize_test.go:15-40.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import ref_bigint as RB

SIZES = [16384, 16384, 256, 16, 65536]


def test_manifest_hashes():
    man = json.load(open(os.path.join(GOLDEN, "MANIFEST.json")))
    for name, meta in man.items():
        data = open(os.path.join(GOLDEN, name), "rb").read()
        assert len(data) == meta["bytes"] and hashlib.sha256(data).hexdigest() == meta["sha256"], name


def test_constants_match_reference_tables(oracle):
    # three in-tree copies (sppark bls12-377.hpp:10-39, ff_config.cuh:13-40, SHM.cu:13-48)
    c = oracle.constants(0)
    assert c["p"] == RB.CURVES[0]["p"] and c["r"] == RB.CURVES[0]["r"]
    assert c["R"] == 0x008d6661e2fdf49a4cf495bf803c84e87b4e97b76e7c63059f7db3a98a7d3ff251409f837fffffb102cdffffffffff68
    assert c["R2"] == 0x006dfccb1e914b88837e92f041790bf9bfdf7d03827dc3ac22a5f11162d6b46d0329fcaab00431b1b786686c9400cd22
    assert c["pinv"] == 0x8508bfffffffffff
    c1 = oracle.constants(1)
    assert c1["p"] == RB.CURVES[1]["p"] and c1["r"] == RB.CURVES[1]["r"] and c1["pinv"] == 0x89f3fffcfffcfffd
    for cid in (0, 1):
        cc, p = oracle.constants(cid), RB.CURVES[cid]["p"]
        assert cc["R"] == (1 << 384) % p and cc["R2"] == (1 << 768) % p
        assert cc["gx"] == RB.CURVES[cid]["gx"] * (1 << 384) % p and cc["b"] == RB.CURVES[cid]["b"] * (1 << 384) % p


def test_ark_window_rule(oracle):
    # N = 2^16 -> 13, 2^20 -> 15, 2^24 -> 18, 2^26 -> 19, 2^28 -> 21; < 32 -> 3
    assert [oracle.window_bits(1 << k) for k in (16, 20, 24, 26, 28)] == [13, 15, 18, 19, 21]
    assert oracle.window_bits(16) == 3 and oracle.window_bits(31) == 3 and oracle.window_bits(33) == 6


@pytest.mark.parametrize("k", range(5))
def test_c_oracle_reproduces_arkworks_results(oracle, golden_instances, k):
    g = golden_instances[k]
    assert len(g["scalars"]) == SIZES[k]
    assert oracle.on_curve(g["aff"])
    jac = oracle.ark_msm(g["aff"], g["scalars"])
    assert oracle.compress(oracle.jac_to_affine(jac)).tobytes() == g["result"]
    if k == 3:
        assert g["result"].hex() == ("e84820ff9ba244d9f3a700a1daf0117b5550fa15c6c75ddc8ad535bc55a2db57"
                                     "056489f85e22762e61f4f6bbd7660680")


@pytest.mark.parametrize("k", [2, 3])
def test_python_bigint_reproduces_arkworks_results(k):
    inst = list(RB.read_instances(os.path.join(GOLDEN, "ark377_points.bin"), os.path.join(GOLDEN, "ark377_scalars.bin")))
    res = open(os.path.join(GOLDEN, "ark377_result.bin"), "rb").read()
    P, S = inst[k]
    pts = [RB.decompress(b) for b in P]
    assert RB.compress(RB.naive_msm(pts, S)) == res[48 * k: 48 * k + 48]


def test_all_infinity_instance(oracle):
    # yyteam/mobileMSM/app/src/main/assets/zero-points-test: 256 bases, every one flagged infinity
    inst = oracle.read_golden(os.path.join(GOLDEN, "ark377_allinf_points.bin"), os.path.join(GOLDEN, "ark377_allinf_scalars.bin"))
    P, S, _ = inst[0]
    aff = oracle.decompress(P)
    assert aff[:, 96].all()
    out = oracle.jac_to_affine(oracle.ark_msm(aff, S))
    assert out[0, 96] == 1


@pytest.mark.parametrize("curve", [0, 1])
def test_c_oracle_vs_python_bigint_random(oracle, curve):
    """Independent implementations agree on fresh inputs (the only pin BLS12-381 has)."""
    n = 24
    aff = oracle.gen_points(n, seed=11 + curve, curve=curve)
    sc = oracle.gen_scalars(n, seed=12 + curve, curve=curve)
    sc[0, :] = 0
    sc[1, :] = 0
    sc[1, 0] = 1                                       # scalar == 1 shortcut
    r = RB.CURVES[curve]["r"]
    sc[2, :] = np.frombuffer((r - 1).to_bytes(32, "little"), np.uint8)
    aff[3, :] = 0
    aff[3, 96] = 1                                     # infinity base
    aff[5] = aff[4]                                    # equal bases
    aff[7] = oracle.aff_neg(aff[6:7], curve)[0]
    sc[7] = sc[6]                                      # P and -P with equal scalars cancel
    assert oracle.on_curve(aff, curve)
    pts = [RB.from_mont_affine_bytes(bytes(a), curve) for a in aff]
    S = [int.from_bytes(bytes(s), "little") for s in sc]
    want = RB.to_mont_affine_bytes(RB.naive_msm(pts, S, curve), curve)
    for c in (0, 5):
        got = oracle.jac_to_affine(oracle.ark_msm(aff, sc, curve, c=c), curve)[0]
        assert bytes(got[:97]) == want[:97]
    got = oracle.jac_to_affine(oracle.naive_msm(aff, sc, curve), curve)[0]
    assert bytes(got[:97]) == want[:97]


def test_pippenger_vs_naive_4096(oracle):
    n = 4096
    aff = oracle.gen_points(64, seed=5)
    aff = np.tile(aff, (n // 64, 1))                   # harness-style replication (P1A/yrrid/src/util.rs:26-28)
    sc = oracle.gen_scalars(n, seed=6)
    a = oracle.jac_to_affine(oracle.ark_msm(aff, sc))
    b = oracle.jac_to_affine(oracle.naive_msm(aff, sc))
    assert a.tobytes() == b.tobytes()
    assert oracle.jac_eq(oracle.ark_msm(aff, sc), oracle.naive_msm(aff, sc))


def test_harness_fixture_files_roundtrip(oracle, tmp_path):
    """TEST_WRITE_DATA_TO / TEST_LOAD_DATA_FROM files of the reference's CPU harness (tests/harness_data.py)."""
    import harness_data as HD
    n, batches = 64, 2
    pts = oracle.gen_points(n, seed=91)
    sc = oracle.gen_scalars(n * batches, seed=92)
    res = np.concatenate([oracle.jac_to_affine(oracle.ark_msm(pts, sc[b * n:(b + 1) * n])) for b in range(batches)])
    HD.write(oracle, str(tmp_path), pts, sc, res)
    assert os.path.getsize(tmp_path / "points.bin") == 8 + 48 * n and os.path.getsize(tmp_path / "scalars.bin") == 8 + 32 * n * batches
    b2, pts2, sc2, res2 = HD.load(oracle, str(tmp_path))
    assert b2 == batches and np.array_equal(pts2, pts) and np.array_equal(sc2, sc) and np.array_equal(res2, res)


def test_bls12_381_small_multiples_of_the_generator(oracle, bls381_small_multiples):
    """BLS12-381 has no MSM known-answer vector in the reference; what it does hold are ten G1 points in the wasm entry's
    group-law tests (open-division/prize4-msm-wasm/Manta-Network/.../test/batchAffine.js:732-745), k * G for
    k = 1, 2, 3, 5, 4, 6, 7, 8, 9, 9.  Both oracles must reproduce every one from the generator, and an MSM over them
    must equal the closed form (sum k_i s_i) G."""
    import known_answer as KA
    aff, ks = bls381_small_multiples
    c = oracle.constants(1)
    assert oracle.on_curve(aff, 1)
    assert int.from_bytes(aff[0, :48].tobytes(), "little") == c["gx"] and int.from_bytes(aff[0, 48:96].tobytes(), "little") == c["gy"]
    for i, k in enumerate(ks):
        assert np.array_equal(KA.expected_affine(oracle, 1, k), aff[i]), k            # C oracle: k * G by double-and-add
        q = RB.mul(k, (RB.CURVES[1]["gx"], RB.CURVES[1]["gy"]), RB.CURVES[1]["p"])      # Python big integers, affine law
        assert RB.to_mont_affine_bytes(q, 1) == aff[i].tobytes(), k
    sc = oracle.gen_scalars(len(ks), seed=381, curve=1)
    t = sum(k * int.from_bytes(sc[i].tobytes(), "little") for i, k in enumerate(ks)) % c["r"]
    exp = KA.expected_affine(oracle, 1, t)
    assert np.array_equal(oracle.jac_to_affine(oracle.ark_msm(aff, sc, 1), 1)[0], exp)
    assert np.array_equal(oracle.jac_to_affine(oracle.naive_msm(aff, sc, 1), 1)[0], exp)


def test_bls12_381_hash_to_curve_vectors_as_two_point_msms(oracle, bls381_hash_to_curve):
    """A BLS12-381 G1 MSM known answer from the reference's own test data: arkworks' RFC 9380 hash-to-curve vectors
    (.../arkworks-algebra/ec/src/hashing/tests/testdata/BLS12381G1_XMD-SHA-256_SSWU_RO_.json, five of them) give Q0, Q1 and
    P = clear_cofactor(Q0 + Q1) = h_eff * Q0 + h_eff * Q1 -- points with full-size coordinates, Q0 and Q1 on the curve but
    OUTSIDE the prime-order subgroup (arkworks' MSM is correct for every on-curve point, short_weierstrass_jacobian.rs:662-729).
    The C restatement of arkworks' Pippenger, the naive sum and the Python big-integer oracle must all return P."""
    vectors, h = bls381_hash_to_curve
    c = oracle.constants(1)
    sc = np.zeros((2, 32), np.uint8)
    sc[:] = np.frombuffer(h.to_bytes(32, "little"), np.uint8)
    total = None
    for bases, want in vectors:
        assert oracle.on_curve(bases, 1) and oracle.on_curve(want[None], 1)
        assert np.array_equal(oracle.jac_to_affine(oracle.ark_msm(bases, sc, 1), 1)[0], want)
        assert np.array_equal(oracle.jac_to_affine(oracle.naive_msm(bases, sc, 1), 1)[0], want)
        q = [RB.from_mont_affine_bytes(b.tobytes(), 1) for b in bases]
        got = RB.add(RB.mul(h, q[0], c["p"]), RB.mul(h, q[1], c["p"]), c["p"])
        assert RB.to_mont_affine_bytes(got, 1) == want.tobytes()
        assert RB.mul(c["r"], got, c["p"]) is None and RB.mul(c["r"], q[0], c["p"]) is not None     # P in G1, Q0 not
        total = RB.add(total, got, c["p"])
    # all ten bases in one MSM, distinct scalars h_eff * k_i (64 + 8 bits): sum k_i (P-part) via the big-integer law
    bases = np.concatenate([b for b, _ in vectors])
    ks = list(range(3, 13))
    sc10 = np.stack([np.frombuffer((h * k).to_bytes(32, "little"), np.uint8) for k in ks])
    want = None
    for (b, _), i in zip(vectors, range(5)):
        for j in range(2):
            want = RB.add(want, RB.mul(h * ks[2 * i + j], RB.from_mont_affine_bytes(b[j].tobytes(), 1), c["p"]), c["p"])
    assert oracle.jac_to_affine(oracle.ark_msm(bases, sc10, 1), 1)[0].tobytes() == RB.to_mont_affine_bytes(want, 1)


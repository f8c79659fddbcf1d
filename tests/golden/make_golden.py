#!/usr/bin/env python3
"""Fixture import script (run in the build container only; /root/reference is absent on the GPU box).

Copies the arkworks-generated BLS12-377 G1 MSM known-answer vectors that the reference ships in its
prize-7 mobile harness into tests/golden/ and records their sha256:

  /root/reference/open-division/prize7-msm-mobile/gnark/msm/bls12-377/{points,scalars,result}
      written by gnark/src/lib.rs:245-283 (result = ark_ec::msm::VariableBaseMSM::multi_scalar_mul,
      ark 0.3, CanonicalSerialize) and read back by gnark/msm/bls12-377/marshal_zprize.go:24-150.
      5 instances, N = 16384, 16384, 256, 16, 65536.
  /root/reference/open-division/prize7-msm-mobile/yyteam/mobileMSM/app/src/main/assets/zero-points-test/
      {points,scalars}: 256 points all flagged infinity (edge case; expected result = infinity).

  /root/reference/open-division/prize4-msm-wasm/Manta-Network/submission/wasmcurves-optimized/test/batchAffine.js
      "g1m add is correct." (l.732-745): ten BLS12-381 G1 points in canonical (non-Montgomery) affine form that the
      reference's wasm tests feed to its group law -- the only BLS12-381 G1 constants the reference holds besides the
      generator.  They are k*G for k = 1, 2, 3, 5, 4, 6, 7, 8, 9, 9 (established below with affine chord-and-tangent
      arithmetic on Python integers, independent of oracle/); written to bls12_381_small_multiples.json.

  /root/reference/open-division/prize4-msm-wasm/snarkify/zprize-prize4-15ac8c55-arkworks-algebra/ec/src/hashing/tests/testdata/
      BLS12381G1_XMD-SHA-256_SSWU_RO_.json (the same file, byte for byte, sits in prize1b-msm-fpga/nickray/driver/algebra and
      prize7-msm-mobile/nistath): the RFC 9380 hash-to-curve vectors arkworks tests its BLS12-381 G1 against.  Each of the
      five gives Q0, Q1 (on the curve, outside the prime-order subgroup) and P = clear_cofactor(Q0 + Q1) =
      h_eff * Q0 + h_eff * Q1 with h_eff = 0xd201000000010001 (RFC 9380 section 8.8.1; arkworks
      bls12_381/src/curves/g1.rs clear_cofactor) -- a two-point MSM with a known answer, on points with full-size coordinates.
      Written to bls12_381_hash_to_curve.json (checked below with the affine law on Python integers, independent of oracle/).

These are DATA fixtures produced by the reference's own oracle (arkworks), not reference sources.
Format: repeat { u64 LE n ; n x 48-byte compressed G1Affine } for points,
repeat { u64 LE n ; n x 32-byte LE canonical scalar } for scalars, one 48-byte compressed point per
instance for result.
"""
import hashlib
import json
import os
import re
import shutil

REF = "/root/reference/open-division/prize7-msm-mobile"
HERE = os.path.dirname(os.path.abspath(__file__))

FILES = {
    "ark377_points.bin": f"{REF}/gnark/msm/bls12-377/points",
    "ark377_scalars.bin": f"{REF}/gnark/msm/bls12-377/scalars",
    "ark377_result.bin": f"{REF}/gnark/msm/bls12-377/result",
    "ark377_allinf_points.bin": f"{REF}/yyteam/mobileMSM/app/src/main/assets/zero-points-test/points",
    "ark377_allinf_scalars.bin": f"{REF}/yyteam/mobileMSM/app/src/main/assets/zero-points-test/scalars",
}


WASM_TEST = "/root/reference/open-division/prize4-msm-wasm/Manta-Network/submission/wasmcurves-optimized/test/batchAffine.js"
P381 = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab


def _affine_add(P, Q, p):
    if P is None or Q is None:
        return P or Q
    (x1, y1), (x2, y2) = P, Q
    if x1 == x2:
        if (y1 + y2) % p == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, p) % p
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
    x3 = (lam * lam - x1 - x2) % p
    return x3, (lam * (x1 - x3) - y1) % p


def import_bls12_381_points():
    src = open(WASM_TEST).read()
    blk = src[src.index('it("g1m add is correct."'):]
    blk = blk[:blk.index("let numPoints")]
    nums = [int(h, 16) for h in re.findall(r"0x([0-9a-f]{80,96})n", blk)]
    assert len(nums) == 20
    pts = [(nums[2 * i], nums[2 * i + 1]) for i in range(10)]
    multiples, acc, seen = [None] * 10, None, {}
    for k in range(1, 33):
        acc = _affine_add(acc, pts[0], P381)
        seen[acc] = k
    multiples = [seen[q] for q in pts]                      # KeyError = not a small multiple of the first point
    assert multiples == [1, 2, 3, 5, 4, 6, 7, 8, 9, 9]
    assert all((y * y - x * x * x - 4) % P381 == 0 for x, y in pts)
    out = {"source": WASM_TEST.replace("/root/reference/", "") + ' ("g1m add is correct.", input points)',
           "form": "canonical affine (x, y), big-endian hex, not Montgomery", "multiples_of_first_point": multiples,
           "points": [["%096x" % x, "%096x" % y] for x, y in pts]}
    with open(os.path.join(HERE, "bls12_381_small_multiples.json"), "w") as f:
        json.dump(out, f, indent=1)
    return out


H2C_TEST = ("/root/reference/open-division/prize4-msm-wasm/snarkify/zprize-prize4-15ac8c55-arkworks-algebra/ec/src/hashing/tests/"
            "testdata/BLS12381G1_XMD-SHA-256_SSWU_RO_.json")
H_EFF_381 = 0xd201000000010001
R381 = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def _affine_mul(k, P, p):
    acc = None
    while k:
        if k & 1:
            acc = _affine_add(acc, P, p)
        P = _affine_add(P, P, p)
        k >>= 1
    return acc


def import_bls12_381_hash_to_curve():
    raw = open(H2C_TEST, "rb").read()
    d = json.loads(raw)
    assert d["curve"] == "BLS12-381 G1" and len(d["vectors"]) == 5
    vectors = []
    for v in d["vectors"]:
        P, Q0, Q1 = [(int(v[k]["x"], 16), int(v[k]["y"], 16)) for k in ("P", "Q0", "Q1")]
        assert all((y * y - x * x * x - 4) % P381 == 0 for x, y in (P, Q0, Q1))
        assert _affine_mul(H_EFF_381, _affine_add(Q0, Q1, P381), P381) == P              # the known answer
        assert _affine_mul(R381, P, P381) is None                                        # P is in G1 ...
        assert _affine_mul(R381, Q0, P381) is not None and _affine_mul(R381, Q1, P381) is not None   # ... Q0, Q1 are not
        vectors.append({"msg": v["msg"], "P": ["%096x" % c for c in P], "Q0": ["%096x" % c for c in Q0], "Q1": ["%096x" % c for c in Q1]})
    out = {"source": H2C_TEST.replace("/root/reference/", ""), "source_sha256": hashlib.sha256(raw).hexdigest(),
           "form": "canonical affine (x, y), big-endian hex, not Montgomery",
           "relation": "P = h_eff * Q0 + h_eff * Q1 (clear_cofactor of Q0 + Q1)", "h_eff": "%x" % H_EFF_381, "vectors": vectors}
    with open(os.path.join(HERE, "bls12_381_hash_to_curve.json"), "w") as f:
        json.dump(out, f, indent=1)
    return out


def main():
    import_bls12_381_points()
    import_bls12_381_hash_to_curve()
    manifest = {}
    for dst, src in FILES.items():
        out = os.path.join(HERE, dst)
        shutil.copyfile(src, out)
        os.chmod(out, 0o644)
        with open(out, "rb") as f:
            data = f.read()
        manifest[dst] = {"source": src.replace("/root/reference/", ""), "bytes": len(data),
                         "sha256": hashlib.sha256(data).hexdigest()}
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print(json.dumps(manifest, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()

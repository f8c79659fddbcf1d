"""The reference harness's fixture mechanism (P1B/hardcaml/zprize/msm_pippenger/test_naive_harness/tests/msm.rs:19-84):
TEST_WRITE_DATA_TO / TEST_LOAD_DATA_FROM name a directory holding ark-serialize files
    points.bin            Vec<G1Affine>   u64 LE count + 48-byte compressed points
    scalars.bin           Vec<Fr>         u64 LE count + 32-byte LE canonical field elements (count = batches * points)
    arkworks_results.bin  Vec<G1Affine>   one compressed point per batch
The harness transmutes the in-memory Fr (Montgomery form) to BigInteger256, so the integer the MSM multiplies by is
s * 2^256 mod r for a stored canonical s (tests/msm.rs:33-36, 90-92); to_msm_scalars / from_msm_scalars convert.
TEST INFRASTRUCTURE: point (de)compression goes through the CPU oracle."""
import os
import struct

import numpy as np


def _r(O):
    return O.constants(0)["r"]


def to_msm_scalars(O, canonical_u8):
    """stored canonical Fr -> the 4 x u64 the harness hands to multi_scalar_mult (Montgomery representation)."""
    r = _r(O)
    out = np.zeros_like(canonical_u8)
    for i, row in enumerate(canonical_u8):
        v = int.from_bytes(bytes(row), "little") * (1 << 256) % r
        out[i] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)
    return out


def from_msm_scalars(O, msm_u8):
    r = _r(O)
    rinv = pow(1 << 256, -1, r)
    out = np.zeros_like(msm_u8)
    for i, row in enumerate(msm_u8):
        v = int.from_bytes(bytes(row), "little") * rinv % r
        out[i] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)
    return out


def write(O, dirname, points_aff, msm_scalars, results_aff):
    os.makedirs(dirname, exist_ok=True)
    with open(os.path.join(dirname, "points.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(points_aff)) + O.compress(points_aff).tobytes())
    sc = from_msm_scalars(O, np.ascontiguousarray(msm_scalars).reshape(-1, 32))
    with open(os.path.join(dirname, "scalars.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(sc)) + sc.tobytes())
    with open(os.path.join(dirname, "arkworks_results.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(results_aff)) + O.compress(results_aff).tobytes())


def load(O, dirname):
    """-> (batches, points uint8[n,104], msm scalars uint8[batches*n,32], results uint8[batches,104])"""
    def vec(name, width):
        raw = open(os.path.join(dirname, name), "rb").read()
        (cnt,) = struct.unpack_from("<Q", raw, 0)
        assert len(raw) == 8 + cnt * width, name
        return cnt, raw[8:]
    n, praw = vec("points.bin", 48)
    m, sraw = vec("scalars.bin", 32)
    b, rraw = vec("arkworks_results.bin", 48)
    assert m % n == 0 and m >= n and b == m // n
    points = O.decompress(praw)
    scalars = to_msm_scalars(O, np.frombuffer(sraw, np.uint8).reshape(m, 32))
    results = O.decompress(rraw)
    return b, points, scalars, results

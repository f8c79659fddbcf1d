"""Independent big-int CPU oracle (TEST INFRASTRUCTURE ONLY; never imported by the product).

Plain Python integers, affine coordinates, schoolbook double-and-add: sum_i s_i * P_i.  It shares no
code and no representation (no Montgomery form, no Jacobian coordinates, no windows) with
oracle/ark_msm_oracle.c, so agreement between the two on the reference's golden vectors
(open-division/prize7-msm-mobile/gnark/msm/bls12-377/{points,scalars,result}) and on
fresh random inputs is a real cross-check of the C restatement that the GPU parity tests rely on.

Reference anchors: constants P1B/nickray/driver/bls12-377/src/fields/fq.rs:4, fr.rs:24,
curves/g1.rs:23,159-165; compressed encoding algebra/serialize/src/flags.rs:64-131 and
algebra/ec/src/models/short_weierstrass_jacobian.rs:806-818; naive MSM
algebra/test-templates/src/msm.rs:7-38.
"""
import struct

CURVES = {
    0: dict(name="bls12-377",
            p=0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001,
            r=0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001, b=1,
            gx=81937999373150964239938255573465948239988671502647976594219695644855304257327692006745978603320413799295628339695,
            gy=241266749859715473739788878240585681733927191168601896383759122102112907357779751001206799952863815012735208165030),
    1: dict(name="bls12-381",
            p=0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
            r=0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001, b=4,
            gx=0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
            gy=0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1),
}
R = 1 << 384


def add(P, Q, p):
    """Affine addition; None is the point at infinity."""
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % p == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, p) % p
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
    x3 = (lam * lam - x1 - x2) % p
    return x3, (lam * (x1 - x3) - y1) % p


def mul(k, P, p):
    acc = None
    while k:
        if k & 1:
            acc = add(acc, P, p)
        P = add(P, P, p)
        k >>= 1
    return acc


def naive_msm(points, scalars, curve=0):
    p = CURVES[curve]["p"]
    acc = None
    for P, s in zip(points, scalars):
        acc = add(acc, mul(s, P, p), p)
    return acc


def sqrt_mod(a, p):
    """Tonelli-Shanks."""
    a %= p
    if a == 0:
        return 0
    if pow(a, (p - 1) // 2, p) != 1:
        return None
    q, s = p - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (p - 1) // 2, p) == 1:
        z += 1
    c, x, t, m = pow(z, q, p), pow(a, (q + 1) // 2, p), pow(a, q, p), s
    while t != 1:
        i, tt = 0, t
        while tt != 1:
            tt = tt * tt % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        x, c, t, m = x * b % p, b * b % p, t * b * b % p, i
    return x


def decompress(buf, curve=0):
    """48-byte ark-serialize compressed G1Affine -> (x, y) canonical ints or None."""
    cv = CURVES[curve]
    p = cv["p"]
    flags = buf[47] & 0xC0
    if flags & 0x40:
        return None
    x = int.from_bytes(buf[:47] + bytes([buf[47] & 0x3F]), "little")
    y = sqrt_mod(x * x * x + cv["b"], p)
    assert y is not None
    larger = y > p - y
    if larger != bool(flags & 0x80):
        y = p - y
    return x, y


def compress(P, curve=0):
    p = CURVES[curve]["p"]
    if P is None:
        return bytes(47) + b"\x40"
    x, y = P
    out = bytearray(x.to_bytes(48, "little"))
    if y > p - y:
        out[47] |= 0x80
    return bytes(out)


def read_instances(points_path, scalars_path):
    """Yield (list of 48-byte compressed points, list of int scalars) per instance."""
    pts, scs = open(points_path, "rb").read(), open(scalars_path, "rb").read()
    po = so = 0
    while po < len(pts):
        (n,) = struct.unpack_from("<Q", pts, po)
        (m,) = struct.unpack_from("<Q", scs, so)
        assert n == m
        P = [pts[po + 8 + 48 * i: po + 8 + 48 * (i + 1)] for i in range(n)]
        S = [int.from_bytes(scs[so + 8 + 32 * i: so + 8 + 32 * (i + 1)], "little") for i in range(n)]
        po += 8 + 48 * n
        so += 8 + 32 * n
        yield P, S


def to_mont_affine_bytes(P, curve=0):
    """(x, y) canonical ints -> 104-byte arkworks G1Affine (Montgomery limbs + infinity flag)."""
    p = CURVES[curve]["p"]
    if P is None:
        return (0).to_bytes(48, "little") + (R % p).to_bytes(48, "little") + b"\x01" + bytes(7)
    x, y = P
    return (x * R % p).to_bytes(48, "little") + (y * R % p).to_bytes(48, "little") + bytes(8)


def from_mont_affine_bytes(buf, curve=0):
    p = CURVES[curve]["p"]
    if buf[96]:
        return None
    ri = pow(R, -1, p)
    return (int.from_bytes(buf[:48], "little") * ri % p, int.from_bytes(buf[48:96], "little") * ri % p)

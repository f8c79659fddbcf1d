// Host build of the device field/curve headers with the PTX carry flag emulated (TEST ONLY).
// Lets the CPU suite check the limb-level algorithms of csrc/fp.cuh and csrc/ec.cuh against the oracle
// without a GPU.  Built by tests/test_host_emul.py with g++; never part of the product library.
#include <cstring>
#include <cstddef>
#include "../../2022-entries_b200/csrc/te.cuh"
using namespace b200msm;

// All entry points take and return arkworks-ABI field elements (48 bytes, R = 2^384, canonical) and convert to the
// device representation (14 x 28-bit limbs, R = 2^392) with the same from_abi / to_abi the kernels use.
template <class C> static void fp_ops(int op, const fp_abi *a, const fp_abi *b, fp_abi *r, size_t n) {
    typedef Fp<C> F;
    for (size_t i = 0; i < n; i++) {
        fp x, y, t;
        F::from_abi(x, a[i]);
        F::from_abi(y, b[i]);
        switch (op) {
        case 0: F::mul(t, x, y); break;
        case 1: F::add(t, x, y); break;
        case 2: F::template sub<2>(t, x, y); break;
        case 3: F::sqr(t, x); break;
        case 4: Ec<C>::inv_ni(t, x); break;         // the windowed inversion the kernels call (ec.cuh)
        case 6: F::mul_karatsuba(t, x, y); break;
        default: t = x;                       // 5: ABI round trip
        }
        F::to_abi(r[i], t);
    }
}
template <class C> static void load_xyzz_abi(xyzz_t &q, const xyzz_abi_t &a) {
    typedef Fp<C> F;
    F::from_abi(q.x, a.x); F::from_abi(q.y, a.y); F::from_abi(q.zz, a.zz); F::from_abi(q.zzz, a.zzz);
}
// op 0: acc += aff (madd); 1: acc += other (add); 2: acc = 2 acc; 3: acc = 2*aff
template <class C> static void ec_ops(int op, xyzz_abi_t *acc, const void *other, size_t n, int reps) {
    typedef Ec<C> E; typedef Fp<C> F;
    for (size_t i = 0; i < n; i++) {
        xyzz_t q; load_xyzz_abi<C>(q, acc[i]);
        affine_t a; xyzz_t o;
        if (op == 0 || op == 3) { F::from_abi(a.x, ((const affine_abi_t *)other)[i].x); F::from_abi(a.y, ((const affine_abi_t *)other)[i].y); }
        if (op == 1) load_xyzz_abi<C>(o, ((const xyzz_abi_t *)other)[i]);
        for (int k = 0; k < reps; k++) {
            switch (op) {
            case 0: E::madd(q, a.x, a.y); break;
            case 1: E::add(q, o); break;
            case 2: E::dbl(q); break;
            case 3: E::dbl_affine(q, a.x, a.y); break;
            }
        }
        F::to_abi(acc[i].x, q.x); F::to_abi(acc[i].y, q.y); F::to_abi(acc[i].zz, q.zz); F::to_abi(acc[i].zzz, q.zzz);
    }
}
// lazy-bound stress: multiply (a + k1 p) by (b + k2 p) -- same residues, larger representatives -- and report
// whether the documented output bound (< 2p) held
template <class C> static int fp_mul_lazy(const fp_abi *a, const fp_abi *b, fp_abi *r, size_t n, int k1, int k2) {
    typedef Fp<C> F;
    int violations = 0;
    for (size_t i = 0; i < n; i++) {
        fp x, y, pp;
        F::from_abi(x, a[i]); F::from_abi(y, b[i]);
        for (int j = 0; j < FP_LIMBS; j++) pp.l[j] = C::p(j);
        for (int k = 0; k < k1; k++) F::add(x, x, pp);
        for (int k = 0; k < k2; k++) F::add(y, y, pp);
        fp t; F::mul(t, x, y);
        fp chk = t; F::csub_p(chk); F::csub_p(chk);          // < 2p  <=>  two conditional subtractions reach [0, p)
        fp canon = t; F::reduce(canon);
        if (memcmp(&chk, &canon, sizeof(fp)) != 0) violations++;
        F::to_abi(r[i], t);
    }
    return violations;
}
// curve constants as plain little-endian integers for the test to re-derive: p, one, cin, cout (4 x 56 bytes of limbs)
template <class C> static void constants(uint32_t *out) {
    for (int i = 0; i < FP_LIMBS; i++) { out[i] = C::p(i); out[14 + i] = C::one(i); out[28 + i] = C::cin(i); out[42 + i] = C::cout(i); }
    out[56] = C::INV; out[57] = C::P0INV;
    for (int k = 0; k < 16; k++) for (int i = 0; i < FP_LIMBS; i++) out[64 + k * 14 + i] = kp_dom<C>(k + 1, i);
}
// twisted-Edwards bucket arithmetic (te.cuh), BLS12-377: a, b short-Weierstrass affine (ABI) -> table entries -> op ->
// back to short-Weierstrass affine.  op 0: a + reps*b (mixed), 1: the same by full adds, 2: 2^reps a, 3: a - reps*b.
// out: x, y (ABI) and a flag word (1 = infinity)
static void te_ops(int op, const affine_abi_t *a, const affine_abi_t *b, uint32_t *out, size_t n, int reps) {
    typedef Bls12_377 C; typedef Fp<C> F; typedef Te<C> T;
    for (size_t i = 0; i < n; i++) {
        te_entry_t ent[2];
        for (int k = 0; k < 2; k++) {
            const affine_abi_t &s = k == 0 ? a[i] : b[i];
            fp x, y, nu, nv, dv, den, inv;
            F::from_abi(x, s.x); F::from_abi(y, s.y);
            T::map_prepare(nu, nv, dv, den, x, y);
            F::inv(inv, den);
            T::map_finish(ent[k], nu, nv, dv, y, inv);
        }
        xyzz_t acc, other; te_pre_t q;
        T::set_inf(acc); T::set_inf(other);
        T::unpack(q, ent[0]); T::madd(acc, q);
        T::unpack(q, ent[1]); T::madd(other, q);
        for (int k = 0; k < reps; k++) {
            if (op == 0) T::madd(acc, q);
            else if (op == 1) T::add(acc, other);
            else if (op == 2) T::dbl(acc);
            else { te_pre_t nq = q; T::negate(nq); T::madd(acc, nq); }
        }
        fp x, y;
        bool finite = T::to_sw_affine(x, y, acc);
        if (!finite) { F::set_zero(x); F::set_one(y); }
        fp_abi ox, oy; F::to_abi(ox, x); F::to_abi(oy, y);
        for (int k = 0; k < 12; k++) { out[i * 26 + k] = ox.w[k]; out[i * 26 + 12 + k] = oy.w[k]; }
        out[i * 26 + 24] = finite ? 0 : 1; out[i * 26 + 25] = 0;
    }
}
template <class C> static void te_constants(uint32_t *out) {
    for (int i = 0; i < FP_LIMBS; i++) { out[i] = C::te_k(i); out[14 + i] = C::te_f(i); out[28 + i] = C::te_s(i); out[42 + i] = C::te_s3(i); out[56 + i] = C::te_cy(i); }
}
// the multiplier on raw limb arrays (14 x u32 each): operands with limbs above 2^28 (the un-normalised sums Te::madd feeds it)
template <class C> static void fp_mul_raw(const uint32_t *a, const uint32_t *b, uint32_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fp x, y, t;
        for (int j = 0; j < FP_LIMBS; j++) { x.l[j] = a[i * FP_LIMBS + j]; y.l[j] = b[i * FP_LIMBS + j]; }
        Fp<C>::mul(t, x, y);
        for (int j = 0; j < FP_LIMBS; j++) r[i * FP_LIMBS + j] = t.l[j];
    }
}
// the 13 x 29-bit field of the accumulate kernel (csrc/fp29.cuh): raw multiply, and the twisted-Edwards mixed add Te29::madd
// driven like te_ops (op 0: a + reps*b, op 3: a - reps*b), entries built and packed exactly as the table builder does
static void fp29_mul_raw(const uint32_t *a, const uint32_t *b, uint32_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fp29 x, y, t;
        for (int j = 0; j < F29_LIMBS; j++) { x.l[j] = a[i * F29_LIMBS + j]; y.l[j] = b[i * F29_LIMBS + j]; }
        Fp29::mul(t, x, y);
        for (int j = 0; j < F29_LIMBS; j++) r[i * F29_LIMBS + j] = t.l[j];
    }
}
static void te29_ops(int op, const affine_abi_t *a, const affine_abi_t *b, uint32_t *out, size_t n, int reps) {
    typedef Bls12_377 C; typedef Fp<C> F; typedef Te<C> T;
    for (size_t i = 0; i < n; i++) {
        te_entry_t ent[2];
        for (int k = 0; k < 2; k++) {
            const affine_abi_t &s = k == 0 ? a[i] : b[i];
            fp x, y, nu, nv, dv, den, inv;
            F::from_abi(x, s.x); F::from_abi(y, s.y);
            T::map_prepare(nu, nv, dv, den, x, y);
            F::inv(inv, den);
            T::map_finish(ent[k], nu, nv, dv, y, inv);
        }
        fp29 v[2][3];
        for (int k = 0; k < 2; k++) for (int c = 0; c < 3; c++) {
            fp_abi t;
            for (int w = 0; w < 12; w++) t.w[w] = ent[k].w[12 * c + w];
            Fp29::repack_from_abi(v[k][c], t);
        }
        te29_t acc;
        Te29::set_inf(acc);
        Te29::madd(acc, v[0][0], v[0][1], v[0][2], false);
        for (int k = 0; k < reps; k++) Te29::madd(acc, v[1][0], v[1][1], v[1][2], op == 3);
        xyzz_t q;
        Te29::export_to(q, acc);
        te29_t back;                                 // round trip of the representation change
        Te29::import_from(back, q);
        Te29::export_to(q, back);
        fp x, y;
        bool finite = T::to_sw_affine(x, y, q);
        if (!finite) { F::set_zero(x); F::set_one(y); }
        fp_abi ox, oy; F::to_abi(ox, x); F::to_abi(oy, y);
        for (int k = 0; k < 12; k++) { out[i * 26 + k] = ox.w[k]; out[i * 26 + 12 + k] = oy.w[k]; }
        out[i * 26 + 24] = finite ? 0 : 1; out[i * 26 + 25] = 0;
    }
}
extern "C" {
int emul_fp29_mul_raw(const uint32_t *a, const uint32_t *b, uint32_t *r, size_t n) { fp29_mul_raw(a, b, r, n); return 0; }
int emul_te29_op(int op, const void *a, const void *b, void *out, size_t n, int reps) { te29_ops(op, (const affine_abi_t *)a, (const affine_abi_t *)b, (uint32_t *)out, n, reps); return 0; }
int emul_fp29_constants(uint32_t *out) {
    for (int i = 0; i < F29_LIMBS; i++) { out[i] = P377x29::p(i); out[13 + i] = P377x29::one(i); out[26 + i] = P377x29::to392(i); }
    for (int k = 0; k < 8; k++) for (int i = 0; i < F29_LIMBS; i++) out[39 + k * 13 + i] = kp_dom29(k + 1, i);
    for (int i = 0; i < FP_LIMBS; i++) { out[160 + i] = Bls12_377::te_c406(i); out[174 + i] = Bls12_377::te_c378(i); }
    return 0;
}
int emul_fp_mul_raw(int curve, const uint32_t *a, const uint32_t *b, uint32_t *r, size_t n) {
    if (curve == 0) fp_mul_raw<Bls12_377>(a, b, r, n); else fp_mul_raw<Bls12_381>(a, b, r, n);
    return 0;
}
int emul_te_op(int op, const void *a, const void *b, void *out, size_t n, int reps) { te_ops(op, (const affine_abi_t *)a, (const affine_abi_t *)b, (uint32_t *)out, n, reps); return 0; }
int emul_te_constants(uint32_t *out) { te_constants<Bls12_377>(out); return 0; }
int emul_fp_op(int curve, int op, const void *a, const void *b, void *r, size_t n) {
    if (curve == 0) fp_ops<Bls12_377>(op, (const fp_abi *)a, (const fp_abi *)b, (fp_abi *)r, n); else fp_ops<Bls12_381>(op, (const fp_abi *)a, (const fp_abi *)b, (fp_abi *)r, n);
    return 0;
}
int emul_ec_op(int curve, int op, void *acc, const void *other, size_t n, int reps) {
    if (curve == 0) ec_ops<Bls12_377>(op, (xyzz_abi_t *)acc, other, n, reps); else ec_ops<Bls12_381>(op, (xyzz_abi_t *)acc, other, n, reps);
    return 0;
}
int emul_fp_mul_lazy(int curve, const void *a, const void *b, void *r, size_t n, int k1, int k2) {
    return curve == 0 ? fp_mul_lazy<Bls12_377>((const fp_abi *)a, (const fp_abi *)b, (fp_abi *)r, n, k1, k2)
                      : fp_mul_lazy<Bls12_381>((const fp_abi *)a, (const fp_abi *)b, (fp_abi *)r, n, k1, k2);
}
int emul_constants(int curve, uint32_t *out) {
    if (curve == 0) constants<Bls12_377>(out); else constants<Bls12_381>(out);
    return 0;
}
}

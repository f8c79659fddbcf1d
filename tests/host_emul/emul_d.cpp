// Host build of csrc/fpd.cuh (the FP64-pipe field) with the FPU rounding toward zero (TEST ONLY).
#include <cfenv>
#include <cstddef>
#include "../../2022-entries_b200/csrc/te.cuh"
using namespace b200msm;
typedef FpD<Bls12_377> F;
struct rz_scope { int old; rz_scope() { old = fegetround(); fesetround(FE_TOWARDZERO); } ~rz_scope() { fesetround(old); } };
extern "C" {
// op 0 mul, 1 add, 2 sub<2>, 3 sqr, 4 inv, 5 round trip
int emuld_fp_op(int op, const void *a_, const void *b_, void *r_, size_t n) {
    rz_scope rz;
    const fp_abi *a = (const fp_abi *)a_, *b = (const fp_abi *)b_; fp_abi *r = (fp_abi *)r_;
    for (size_t i = 0; i < n; i++) {
        fpd x, y, t;
        F::from_abi(x, a[i]); F::from_abi(y, b[i]);
        switch (op) {
        case 0: F::mul(t, x, y); break;
        case 1: F::add(t, x, y); break;
        case 2: F::sub<2>(t, x, y); break;
        case 3: F::sqr(t, x); break;
        case 4: F::inv(t, x); break;
        default: t = x;
        }
        F::to_abi(r[i], t);
    }
    return 0;
}
// (a + k1 p)(b + k2 p) with un-normalised limbs: returns how many results were not below 2p
int emuld_fp_mul_lazy(const void *a_, const void *b_, void *r_, size_t n, int k1, int k2) {
    rz_scope rz;
    const fp_abi *a = (const fp_abi *)a_, *b = (const fp_abi *)b_; fp_abi *r = (fp_abi *)r_;
    int bad = 0;
    for (size_t i = 0; i < n; i++) {
        fpd x, y, pp, t;
        F::from_abi(x, a[i]); F::from_abi(y, b[i]);
        for (int j = 0; j < FD_LIMBS; j++) pp.l[j] = (double)fd_p<Bls12_377>(j);
        for (int k = 0; k < k1; k++) { F::add(x, x, pp); if ((k & 1) == 1) F::norm(x); }
        for (int k = 0; k < k2; k++) { F::add(y, y, pp); if ((k & 1) == 1) F::norm(y); }
        F::mul(t, x, y);
        fpd chk = t; F::csub_p(chk); F::csub_p(chk);
        fpd canon = t; F::reduce(canon);
        for (int j = 0; j < FD_LIMBS; j++) if (chk.l[j] != canon.l[j]) { bad++; break; }
        for (int j = 0; j < FD_LIMBS - 1; j++) if (t.l[j] >= 0x1p48) { bad++; break; }
        F::to_abi(r[i], t);
    }
    return bad;
}
int emuld_is_zero(const void *a_, int k, int sub_form) {
    rz_scope rz;
    fpd x, pp; F::from_abi(x, *(const fp_abi *)a_);
    for (int j = 0; j < FD_LIMBS; j++) pp.l[j] = (double)fd_p<Bls12_377>(j);
    if (sub_form) { fpd y = x; F::sub<4>(x, x, y); }      // x - x + 4p: limb-dominant representation of 4p
    for (int i = 0; i < k; i++) { F::add(x, x, pp); if (i & 1) F::norm(x); }
    return F::is_zero(x) ? 1 : 0;
}
// FP64-pipe twisted-Edwards accumulator (te.cuh TeD): entries built by the integer map, a + reps*(+-b) accumulated in
// the FP64 representation starting from the neutral element, exported through to_int, mapped back to short Weierstrass.
int emuld_te_op(int op, const void *a_, const void *b_, void *out_, size_t n, int reps) {
    rz_scope rz;
    typedef Bls12_377 C; typedef Fp<C> FI; typedef Te<C> T; typedef TeD<C> D;
    const affine_abi_t *a = (const affine_abi_t *)a_, *b = (const affine_abi_t *)b_; uint32_t *out = (uint32_t *)out_;
    for (size_t i = 0; i < n; i++) {
        te_entry_t ent[2];
        for (int k = 0; k < 2; k++) {
            const affine_abi_t &s = k == 0 ? a[i] : b[i];
            fp x, y, nu, nv, dv, den, inv;
            FI::from_abi(x, s.x); FI::from_abi(y, s.y);
            T::map_prepare(nu, nv, dv, den, x, y);
            FI::inv(inv, den);
            T::map_finish(ent[k], nu, nv, dv, y, inv, true);       // arkworks-form entries: what the FP64 accumulator reads
        }
        ted_t acc; ted_pre_t q;
        D::set_inf(acc);
        D::unpack(q, ent[0]); D::madd(acc, q);
        D::unpack(q, ent[1]);
        if (op == 3) D::negate(q);
        for (int k = 0; k < reps; k++) D::madd(acc, q);
        xyzz_t r; D::to_int(r, acc);
        fp x, y;
        bool finite = T::to_sw_affine(x, y, r);
        if (!finite) { FI::set_zero(x); FI::set_one(y); }
        fp_abi ox, oy; FI::to_abi(ox, x); FI::to_abi(oy, y);
        for (int k = 0; k < 12; k++) { out[i * 26 + k] = ox.w[k]; out[i * 26 + 12 + k] = oy.w[k]; }
        out[i * 26 + 24] = finite ? 0 : 1; out[i * 26 + 25] = 0;
    }
    return 0;
}
void emuld_constants(uint64_t *out) {
    for (int i = 0; i < 8; i++) { out[i] = fd_p<Bls12_377>(i); out[8 + i] = fd_one<Bls12_377>(i); }
    out[16] = fd_ninv<Bls12_377>(); out[17] = fd_pinv<Bls12_377>();
    for (int k = 0; k < 16; k++) for (int i = 0; i < 8; i++) out[24 + 8 * k + i] = fd_kp_dom<Bls12_377>(k + 1, i);
}
}

// Host build of csrc/fpd.cuh (the FP64-pipe field) with the FPU rounding toward zero (TEST ONLY).
#include <cfenv>
#include <cstddef>
#include "../../2022-entries_b200/csrc/fpd.cuh"
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
void emuld_constants(uint64_t *out) {
    for (int i = 0; i < 8; i++) { out[i] = fd_p<Bls12_377>(i); out[8 + i] = fd_one<Bls12_377>(i); }
    out[16] = fd_ninv<Bls12_377>(); out[17] = fd_pinv<Bls12_377>();
    for (int k = 0; k < 16; k++) for (int i = 0; i < 8; i++) out[24 + 8 * k + i] = fd_kp_dom<Bls12_377>(k + 1, i);
}
}

// ec.cuh -- short-Weierstrass (a = 0) group law in XYZZ ("extended Jacobian") coordinates:
// x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; infinity is ZZ == 0.
//
// Same coordinate system and formulas (EFD madd-2008-s, add-2008-s, dbl-2008-s-1, mdbl-2008-s-1) the reference
// entries accumulate buckets in (sppark xyzz_t.hpp:97-249, matter-labs ec.cuh:418-598, yrrid Curve.cu:630-1127);
// the code below is written from the formulas, with its own lazy-reduction bookkeeping on top of fp.cuh:
//
//   every field value is kept normalised (28-bit limbs) but only lazily reduced: stored coordinates obey
//   X < 10p, Y < 5p, ZZ, ZZZ < 2p.  Fp::sub<K>(a, b) = a + K p - b needs value(b) < (K-1) p, and every multiply
//   needs (bound a)(bound b) < 2^11 (BLS12-381; 2^15 on BLS12-377) for its result to be < 2p -- the largest product
//   below is 13 * 13.  The per-line comments give the bound of each result in units of p.
//
// The exceptional cases of the group law are all handled (the harness replicates 2^15 distinct bases to 2^26, so
// P+P and P+(-P) inside one bucket do occur): accumulator at infinity, equal points
// (doubling) and opposite points (result infinity).
#pragma once
#include "fp.cuh"

#if defined(__CUDACC__)
#define EC_NI __device__ __noinline__
#else
#define EC_NI inline
#endif

namespace b200msm {

struct affine_t { fp x, y; };            // canonical (< p) Montgomery limbs; 112 bytes in the device tables
struct xyzz_t { fp x, y, zz, zzz; };     // 224 bytes
struct affine_abi_t { fp_abi x, y; };    // 96 bytes: arkworks (x, y), R = 2^384
struct xyzz_abi_t { fp_abi x, y, zz, zzz; };

// EC_MUL_BY_CALL (defined by the hot translation units): field multiply / square become real calls with operands and
// result in registers (by-value ABI).  One copy of each body per kernel instead of ten inlined per mixed add keeps
// the accumulate loop inside the instruction cache (ncu: stall_no_instruction 2.5 cycles per issue when inlined).
#if defined(__CUDACC__) && defined(EC_MUL_BY_CALL)
template <class C> __device__ __noinline__ fp fp_mul_call(fp a, fp b) { fp r; Fp<C>::mul(r, a, b); return r; }
template <class C> __device__ __noinline__ fp fp_sqr_call(fp a) { fp r; Fp<C>::sqr(r, a); return r; }
#endif

template <class C> struct Ec {
    typedef Fp<C> F;
#if defined(__CUDACC__) && defined(EC_MUL_BY_CALL)
    __device__ __forceinline__ static void M(fp &r, const fp &a, const fp &b) { r = fp_mul_call<C>(a, b); }
    __device__ __forceinline__ static void S(fp &r, const fp &a) { r = fp_sqr_call<C>(a); }
#else
    FP_HD static void M(fp &r, const fp &a, const fp &b) { F::mul(r, a, b); }
    FP_HD static void S(fp &r, const fp &a) { F::sqr(r, a); }
#endif

    FP_HD static void set_inf(xyzz_t &a) {
        F::set_zero(a.x); F::set_zero(a.y); F::set_zero(a.zz); F::set_zero(a.zzz);
    }
    FP_HD static bool is_inf(const xyzz_t &a) { return F::is_zero(a.zz); }
    FP_HD static void from_affine(xyzz_t &a, const fp &x, const fp &y) {
        a.x = x; a.y = y; F::set_one(a.zz); F::set_one(a.zzz);
    }
    // 2 * (x, y) for an affine point (mdbl-2008-s-1, a = 0); x, y < 2p
    FP_HD static void dbl_affine(xyzz_t &r, const fp &x, const fp &y) {
        fp u, v, w, s, m, t;
        F::add(u, y, y);                  // U = 2Y                    4
        S(v, u);                     // V = U^2                   2
        M(w, u, v);                  // W = U*V                   2
        M(s, x, v);                  // S = X*V                   2
        S(t, x);
        F::add(m, t, t);
        F::add(m, m, t);                  // M = 3X^2                  6
        S(r.x, m);
        F::add(t, s, s);                  // 2S                        4
        F::template sub<5>(r.x, r.x, t);  // X3 = M^2 - 2S             7
        F::template sub<8>(t, s, r.x);    // S - X3                    10
        M(t, m, t);                  // 6 * 10
        M(s, w, y);
        F::template sub<3>(r.y, t, s);    // Y3                        5
        r.zz = v;
        r.zzz = w;
    }
    // a = 2a (dbl-2008-s-1, a = 0).  Infinity stays infinity (ZZ3 = V*0 = 0).
    FP_HD static void dbl(xyzz_t &a) {
        fp u, v, w, s, m, t;
        F::add(u, a.y, a.y);              // U = 2Y1                   10
        S(v, u);                     // 100
        M(w, u, v);                  // 10 * 2
        M(s, a.x, v);                // 10 * 2
        S(t, a.x);                   // 100
        F::add(m, t, t);
        F::add(m, m, t);                  // M = 3 X1^2                6
        fp x3, y3;
        S(x3, m);                    // 36
        F::add(t, s, s);                  //                           4
        F::template sub<5>(x3, x3, t);    // X3                        7
        F::template sub<8>(t, s, x3);     // S - X3                    10
        M(t, m, t);                  // 60
        M(y3, w, a.y);               // 2 * 5
        F::template sub<3>(a.y, t, y3);   // Y3                        5
        a.x = x3;
        M(a.zz, v, a.zz);
        M(a.zzz, w, a.zzz);
    }
    // acc += (x2, y2) with (x2, y2) affine, < 2p, not infinity  (madd-2008-s: 8M + 2S)
    FP_HD static void madd(xyzz_t &acc, const fp &x2, const fp &y2) {
        if (is_inf(acc)) { from_affine(acc, x2, y2); return; }
        fp p, r, pp, ppp, q, t;
        M(p, x2, acc.zz);                  // U2                       2
        M(r, y2, acc.zzz);                 // S2                       2
        F::template sub<11>(p, p, acc.x);       // P = U2 - X1              13
        F::template sub<6>(r, r, acc.y);        // R = S2 - Y1              8
        if (F::is_zero(p)) {
            if (F::is_zero(r)) dbl_affine(acc, x2, y2); else set_inf(acc);
            return;
        }
        S(pp, p);                          // 13 * 13
        M(ppp, p, pp);                     // 13 * 2
        M(q, acc.x, pp);                   // 10 * 2
        S(t, r);                           // 8 * 8
        F::template sub<3>(t, t, ppp);          // R^2 - PPP                5
        fp q2;
        F::add(q2, q, q);                       // 2Q                       4
        F::template sub<5>(t, t, q2);           // X3                       10
        F::template sub<11>(q, q, t);           // Q - X3                   13
        M(q, r, q);                        // 8 * 13
        M(r, acc.y, ppp);                  // 5 * 2
        F::template sub<3>(acc.y, q, r);        // Y3                       5
        acc.x = t;
        M(acc.zz, acc.zz, pp);
        M(acc.zzz, acc.zzz, ppp);
    }
    // a += b  (add-2008-s: 12M + 2S)
    FP_HD static void add(xyzz_t &a, const xyzz_t &b) {
        if (is_inf(b)) return;
        if (is_inf(a)) { a = b; return; }
        fp u1, u2, s1, s2, pp, ppp, q, t;
        M(u1, a.x, b.zz);                  // 10 * 2
        M(u2, b.x, a.zz);
        M(s1, a.y, b.zzz);                 // 5 * 2
        M(s2, b.y, a.zzz);
        F::template sub<3>(u2, u2, u1);         // P                        5
        F::template sub<3>(s2, s2, s1);         // R                        5
        if (F::is_zero(u2)) {
            if (F::is_zero(s2)) dbl(a); else set_inf(a);
            return;
        }
        S(pp, u2);                         // 25
        M(ppp, u2, pp);
        M(q, u1, pp);
        S(t, s2);
        F::template sub<3>(t, t, ppp);          //                          5
        fp q2;
        F::add(q2, q, q);                       //                          4
        F::template sub<5>(t, t, q2);           // X3                       10
        F::template sub<11>(q, q, t);           // Q - X3                   13
        M(q, s2, q);                       // 5 * 13
        M(s1, s1, ppp);
        F::template sub<3>(a.y, q, s1);         // Y3                       5
        a.x = t;
        M(a.zz, a.zz, b.zz);
        M(a.zz, a.zz, pp);
        M(a.zzz, a.zzz, b.zzz);
        M(a.zzz, a.zzz, ppp);
    }
    // canonical affine (x, y) of a finite XYZZ point:  x = X * ZZ^2 / ZZZ^2,  y = Y / ZZZ   (one inversion)
    FP_HD static void to_affine(fp &x, fp &y, const xyzz_t &a) {
        fp i, i2, z2;
        fp zzz = a.zzz;
        F::inv(i, zzz);
        F::sqr(i2, i);
        F::sqr(z2, a.zz);
        F::mul(x, a.x, z2);
        F::mul(x, x, i2);
        F::mul(y, a.y, i);
        F::csub_p(x);
        F::csub_p(y);
    }
    // Non-inlined entry points for the cold kernels (bucket reduction, finalize, table build): one copy of each body
    // per curve instead of one per call site.
    EC_NI static void add_ni(xyzz_t &a, const xyzz_t &b) { add(a, b); }
    EC_NI static void dbl_ni(xyzz_t &a) { dbl(a); }
    EC_NI static void madd_ni(xyzz_t &acc, const fp &x2, const fp &y2) { madd(acc, x2, y2); }
    EC_NI static void dbl_affine_ni(xyzz_t &r, const fp &x, const fp &y) { dbl_affine(r, x, y); }
    EC_NI static void mul_ni(fp &r, const fp &a, const fp &b) { F::mul(r, a, b); }
    EC_NI static void sqr_ni(fp &r, const fp &a) { F::sqr(r, a); }
    // a^(p-2) with a 4-bit fixed window and real squarings: ~376 squarings + ~95 + 14 multiplies instead of ~570 multiplies for
    // the bit-by-bit ladder.  One thread's inversion is on the critical path of every MSM (export of the result) and of every
    // batch of table entries.  28-bit limbs: a window never straddles two limbs.
    EC_NI static void inv_ni(fp &r, const fp &a) {
        fp tbl[16];
        F::set_one(tbl[0]);
        tbl[1] = a;
        for (int w = 2; w < 16; w++) mul_ni(tbl[w], tbl[w - 1], a);
        fp acc;
        F::set_one(acc);
        bool started = false;
        for (int i = FP_LIMBS * FP_BITS - 4; i >= 0; i -= 4) {
            const uint32_t w = (F::pm2_limb(i / FP_BITS) >> (i % FP_BITS)) & 15u;
            if (started) { sqr_ni(acc, acc); sqr_ni(acc, acc); sqr_ni(acc, acc); sqr_ni(acc, acc); }
            if (w) {
                if (started) mul_ni(acc, acc, tbl[w]);
                else { acc = tbl[w]; started = true; }
            }
        }
        r = acc;
    }
    EC_NI static void from_abi_ni(fp &r, const fp_abi &a) { F::from_abi(r, a); }
    EC_NI static void to_abi_ni(fp_abi &r, const fp &a) { F::to_abi(r, a); }
    EC_NI static void to_affine_ni(fp &x, fp &y, const xyzz_t &a) {
        fp i, i2, z2;
        inv_ni(i, a.zzz);
        mul_ni(i2, i, i);
        mul_ni(z2, a.zz, a.zz);
        mul_ni(x, a.x, z2);
        mul_ni(x, x, i2);
        mul_ni(y, a.y, i);
        F::csub_p(x);
        F::csub_p(y);
    }
};

}  // namespace b200msm

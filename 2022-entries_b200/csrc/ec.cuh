// ec.cuh -- short-Weierstrass (a = 0) group law in XYZZ ("extended Jacobian") coordinates:
// x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; infinity is ZZ == 0.
//
// Same coordinate system and formulas (EFD madd-2008-s, add-2008-s, dbl-2008-s-1, mdbl-2008-s-1) the reference
// entries accumulate buckets in (sppark xyzz_t.hpp:97-249, matter-labs ec.cuh:418-598, yrrid Curve.cu:630-1127);
// the code below is written from the formulas, with its own lazy-reduction bookkeeping:
//
//   BLS12-377 (Fp::subx<K> = a + K*p - b, no reduction): stored coordinates obey  X < 8p, Y < 4p, ZZ, ZZZ < 2p;
//   every Montgomery multiply below has (bound a)*(bound b) <= 100 <= SLACK = 152 so its result is < 2p.
//   BLS12-381 (C::TIGHT): every add/sub result is conditionally reduced below 2p, so all operands are < 2p.
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

struct affine_t { fp x, y; };            // canonical (< p) Montgomery limbs; 96 bytes in the device tables
struct xyzz_t { fp x, y, zz, zzz; };     // 192 bytes

template <class C> struct Ec {
    typedef Fp<C> F;

    FP_HD static void set_inf(xyzz_t &a) {
        F::set_zero(a.x); F::set_zero(a.y); F::set_zero(a.zz); F::set_zero(a.zzz);
    }
    FP_HD static bool is_inf(const xyzz_t &a) { return F::is_zero(a.zz); }
    FP_HD static void from_affine(xyzz_t &a, const fp &x, const fp &y) {
        a.x = x; a.y = y; F::set_one(a.zz); F::set_one(a.zzz);
    }
    // 2 * (x, y) for an affine point (mdbl-2008-s-1, a = 0): 4M + 3S... here 3M + 3S + small
    FP_HD static void dbl_affine(xyzz_t &r, const fp &x, const fp &y) {
        fp u, v, w, s, m, t;
        F::addx(u, y, y);                 // U = 2Y            < 2p
        F::sqr(v, u);                     // V = U^2
        F::mul(w, u, v);                  // W = U*V
        F::mul(s, x, v);                  // S = X*V
        F::sqr(t, x);
        F::addx(m, t, t);
        F::addx(m, m, t);                 // M = 3X^2          < 6p
        F::sqr(r.x, m);
        F::addx(t, s, s);                 // 2S                < 4p
        F::template subx<4>(r.x, r.x, t); // X3 = M^2 - 2S     < 6p
        F::template subx<8>(t, s, r.x);   // S - X3            < 10p
        F::mul(t, m, t);                  // 6 * 10 = 60
        F::mul(s, w, y);
        F::template subx<2>(r.y, t, s);   // Y3                < 4p
        r.zz = v;
        r.zzz = w;
    }
    // a = 2a (dbl-2008-s-1, a = 0).  Infinity stays infinity (ZZ3 = V*0 = 0).
    FP_HD static void dbl(xyzz_t &a) {
        fp u, v, w, s, m, t;
        F::addx(u, a.y, a.y);             // U = 2Y1           < 8p
        F::sqr(v, u);                     // 64
        F::mul(w, u, v);                  // 8 * 2
        F::mul(s, a.x, v);                // 8 * 2
        F::sqr(t, a.x);                   // 64
        F::addx(m, t, t);
        F::addx(m, m, t);                 // M = 3 X1^2        < 6p
        fp x3, y3;
        F::sqr(x3, m);                    // 36
        F::addx(t, s, s);
        F::template subx<4>(x3, x3, t);   // < 6p
        F::template subx<8>(t, s, x3);    // < 10p
        F::mul(t, m, t);                  // 60
        F::mul(y3, w, a.y);               // 2 * 4
        F::template subx<2>(a.y, t, y3);  // < 4p
        a.x = x3;
        F::mul(a.zz, v, a.zz);
        F::mul(a.zzz, w, a.zzz);
    }
    // acc += (x2, y2) with (x2, y2) affine, canonical, not infinity  (madd-2008-s: 8M + 2S)
    FP_HD static void madd(xyzz_t &acc, const fp &x2, const fp &y2) {
        if (is_inf(acc)) { from_affine(acc, x2, y2); return; }
        fp p, r, pp, ppp, q, t;
        F::mul(p, x2, acc.zz);                  // U2                       < 2p
        F::mul(r, y2, acc.zzz);                 // S2
        F::template subx<8>(p, p, acc.x);       // P = U2 - X1              < 10p
        F::template subx<4>(r, r, acc.y);       // R = S2 - Y1              < 6p
        if (F::is_zero(p)) {
            if (F::is_zero(r)) dbl_affine(acc, x2, y2); else set_inf(acc);
            return;
        }
        F::sqr(pp, p);                          // 100
        F::mul(ppp, p, pp);                     // 10 * 2
        F::mul(q, acc.x, pp);                   // 8 * 2
        F::sqr(t, r);                           // 36
        F::template subx<2>(t, t, ppp);         // R^2 - PPP                < 4p
        fp q2;
        F::addx(q2, q, q);                      // 2Q                       < 4p
        F::template subx<4>(t, t, q2);          // X3                       < 8p
        F::template subx<8>(q, q, t);           // Q - X3                   < 10p
        F::mul(q, r, q);                        // 6 * 10
        F::mul(r, acc.y, ppp);                  // 4 * 2
        F::template subx<2>(acc.y, q, r);       // Y3                       < 4p
        acc.x = t;
        F::mul(acc.zz, acc.zz, pp);
        F::mul(acc.zzz, acc.zzz, ppp);
    }
    // a += b  (add-2008-s: 12M + 2S)
    FP_HD static void add(xyzz_t &a, const xyzz_t &b) {
        if (is_inf(b)) return;
        if (is_inf(a)) { a = b; return; }
        fp u1, u2, s1, s2, pp, ppp, q, t;
        F::mul(u1, a.x, b.zz);                  // 8 * 2
        F::mul(u2, b.x, a.zz);
        F::mul(s1, a.y, b.zzz);                 // 4 * 2
        F::mul(s2, b.y, a.zzz);
        F::template subx<2>(u2, u2, u1);        // P                        < 4p
        F::template subx<2>(s2, s2, s1);        // R                        < 4p
        if (F::is_zero(u2)) {
            if (F::is_zero(s2)) dbl(a); else set_inf(a);
            return;
        }
        F::sqr(pp, u2);                         // 16
        F::mul(ppp, u2, pp);
        F::mul(q, u1, pp);
        F::sqr(t, s2);
        F::template subx<2>(t, t, ppp);         // < 4p
        fp q2;
        F::addx(q2, q, q);                      // < 4p
        F::template subx<4>(t, t, q2);          // X3 < 8p
        F::template subx<8>(q, q, t);           // < 10p
        F::mul(q, s2, q);                       // 4 * 10
        F::mul(s1, s1, ppp);
        F::template subx<2>(a.y, q, s1);        // Y3 < 4p
        a.x = t;
        F::mul(a.zz, a.zz, b.zz);
        F::mul(a.zz, a.zz, pp);
        F::mul(a.zzz, a.zzz, b.zzz);
        F::mul(a.zzz, a.zzz, ppp);
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
        F::reduce(x);
        F::reduce(y);
    }
    // Non-inlined entry points for the cold kernels (bucket reduction, finalize, table build): one copy of each body
    // per curve instead of one per call site.
    EC_NI static void add_ni(xyzz_t &a, const xyzz_t &b) { add(a, b); }
    EC_NI static void dbl_ni(xyzz_t &a) { dbl(a); }
    EC_NI static void madd_ni(xyzz_t &acc, const fp &x2, const fp &y2) { madd(acc, x2, y2); }
    EC_NI static void dbl_affine_ni(xyzz_t &r, const fp &x, const fp &y) { dbl_affine(r, x, y); }
    EC_NI static void mul_ni(fp &r, const fp &a, const fp &b) { F::mul(r, a, b); }
    EC_NI static void inv_ni(fp &r, const fp &a) {
        fp acc;
        F::set_one(acc);
        for (int i = 383; i >= 0; i--) {
            mul_ni(acc, acc, acc);
            if ((F::pm2_limb(i >> 5) >> (i & 31)) & 1) mul_ni(acc, acc, a);
        }
        r = acc;
    }
    EC_NI static void to_affine_ni(fp &x, fp &y, const xyzz_t &a) {
        fp i, i2, z2;
        inv_ni(i, a.zzz);
        mul_ni(i2, i, i);
        mul_ni(z2, a.zz, a.zz);
        mul_ni(x, a.x, z2);
        mul_ni(x, x, i2);
        mul_ni(y, a.y, i);
        F::reduce(x);
        F::reduce(y);
    }
};

}  // namespace b200msm

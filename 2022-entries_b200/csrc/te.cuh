// te.cuh -- bucket arithmetic in the twisted-Edwards model of the curve (BLS12-377 G1 has one; BLS12-381 does not).
//
//   E_te:  -x^2 + y^2 = 1 + d x^2 y^2,  extended coordinates (X : Y : Z : T), x = X/Z, y = Y/Z, T = XY/Z.
//
// Why: the accumulate kernel is bound by the integer-multiply pipe (fmaheavy 88-89 % busy, profiles/
// r01_ncu_accumulate_v3_calls.txt, r01_ncu_accumulate_v5_te_karatsuba.txt), so the only way down is fewer multiplies per point.  With the base tables stored as
// (y - x, y + x, 2 d x y) the unified mixed addition add-2008-hwcd-3 (a = -1) costs 7 multiplies against 8 M + 2 S for
// XYZZ, and it is the SAME formula for P + Q, P + P, P + (-P) and P + 0: no infinity / doubling branches in the hot
// loop.  d is a square here, so the law is not complete on the whole curve, but its exceptional pairs differ by a
// point of even order; bases in G1 (prime order r) never meet one -- the same contract the reference's entries have
// (matter-labs asserts !infinity and G1 membership is arkworks' deserialisation check).
// The reference's Trapdoor-Tech entry uses the same model (P1A/Trapdoor-Tech/sppark/ec/exte_t.hpp:74-200,
// msm_opt.md); the map below is derived from the curve equation (SW -> Montgomery -> TE, then x scaled so a = -1),
// with arkworks' parameter values (P1B/nickray/driver/bls12-377/src/curves/g1.rs:52-113) as the cross-check.
//
//   SW (x, y), y^2 = x^3 + 1   ->   x' = f (x + 1) / y,   y' = (s (x + 1) - 1) / (s (x + 1) + 1)         s = 1/sqrt 3
//   TE (X : Y : Z)             ->   x = sqrt3 (Z + Y) / (Z - Y) - 1,   y = (f / s) Z (Z + Y) / (X (Z - Y))
//
// Storage: a bucket is 4 field elements like XYZZ (struct xyzz_t reused: x = X, y = Y, zz = Z, zzz = T); a table entry
// is the three precomputed values, canonical, bit-packed 12 x 32 each = 144 bytes (9 x 16: cp.async friendly).
// Value bounds (units of p): every stored coordinate is a multiply output (< 2); the largest product below is 7 x 6.
#pragma once
#include "ec.cuh"
#include "fpd.cuh"
#include "fp29.cuh"

namespace b200msm {

#if defined(TE_ENTRY_160)
struct te_entry_t { uint32_t w[36]; uint32_t pad[4]; };      // experiment: 160-byte, 32-byte-aligned entries = exactly 5 DRAM sectors (profiles/r02_entry160.txt)
#else
struct te_entry_t { uint32_t w[36]; };       // (y - x, y + x, 2 d x y), each a canonical value bit-packed 12 x 32.  Montgomery
                                             // form: 2^406 for the integer accumulator (fp29.cuh: unpacking in the hot loop is
                                             // bit slicing), 2^384 (arkworks') when the context runs the FP64-pipe accumulator
#endif
struct te_pre_t { fp ymx, ypx, kt; };        // the same in registers

template <class C> struct Te {
    typedef Fp<C> F;
    typedef Ec<C> E;

    FP_HD static void load_const(fp &r, int which) {
#pragma unroll
        for (int i = 0; i < FP_LIMBS; i++)
            r.l[i] = which == 0 ? C::te_k(i) : which == 1 ? C::te_f(i) : which == 2 ? C::te_s(i) : which == 3 ? C::te_s3(i) : which == 4 ? C::te_cy(i)
                   : which == 5 ? C::te_c406(i) : C::te_c378(i);
    }
    FP_HD static void set_inf(xyzz_t &a) { F::set_zero(a.x); F::set_one(a.y); F::set_one(a.zz); F::set_zero(a.zzz); }
    // the neutral element is the only point of G1 with X = 0
    FP_HD static bool is_inf(const xyzz_t &a) { return F::is_zero(a.x); }

    // The cold kernels' view of an entry (this file's 14 x 28 arithmetic, Montgomery 2^392): one multiply per value either
    // way -- ark = false: the entry holds 2^406-form values (the hot loop's form); ark = true: arkworks form (2^384)
    FP_HD static void unpack_one(fp &r, const uint32_t *w, bool ark) {
        fp_abi t;
#pragma unroll
        for (int i = 0; i < 12; i++) t.w[i] = w[i];
        if (ark) { F::from_abi(r, t); return; }
        fp v, c;
        F::repack_from_abi(v, t);
        load_const(c, 6);
        F::mul(r, v, c);                          // x 2^406 * 2^378 / 2^392 = x 2^392
    }
    FP_HD static void unpack(te_pre_t &q, const te_entry_t &e, bool ark = false) {
        unpack_one(q.ymx, e.w, ark); unpack_one(q.ypx, e.w + 12, ark); unpack_one(q.kt, e.w + 24, ark);
    }
    FP_HD static void pack_one(uint32_t *w, fp v, bool ark) {
        fp_abi t;
        if (ark) F::to_abi(t, v);
        else {
            fp c;
            load_const(c, 5);
            F::norm(v);
            F::mul(v, v, c);                      // x 2^392 * 2^406 / 2^392 = x 2^406
            F::csub_p(v);
            F::repack_to_abi(t, v);
        }
#pragma unroll
        for (int i = 0; i < 12; i++) w[i] = t.w[i];
    }
    // canonicalise (< p) and pack
    FP_HD static void pack(te_entry_t &e, const te_pre_t &q, bool ark = false) {
        pack_one(e.w, q.ymx, ark); pack_one(e.w + 12, q.ypx, ark); pack_one(e.w + 24, q.kt, ark);
    }
    // -(x, y) = (-x, y): swap the first two values, negate the third
    FP_HD static void negate(te_pre_t &q) {
        fp t = q.ymx; q.ymx = q.ypx; q.ypx = t;
        F::template neg<2>(q.kt, q.kt);
    }

    // acc += q   (add-2008-hwcd-3 with Z2 = 1 and k T2 precomputed: 7 M)
#if defined(TE_LAZY_NORM)
    // The same with the carry propagation left out wherever the multiplier's 64-bit columns allow it.  Notation: L = bound
    // of limbs 0..12 in units of 2^28 (the top limb is bounded by the value and stays below 2^17), V = value in units of p.
    // Stored coordinates and entry values are multiplier outputs or canonical: L 1, V 2.  Karatsuba (Fp::mul_karatsuba)
    // with operands of limb bounds La, Lb: half sums 2 La, 2 Lb; a middle column holds 7 products: 28 La Lb 2^56; the
    // z0 / z2 columns 7 La Lb; after z1 is folded in a column is below (28 + 7) La Lb 2^56, the reduction adds 14 more
    // products m p_j < 2^56 and a carry < 2^36:  (35 La Lb + 15) 2^56 < 2^64  <=>  La Lb <= 6.
    FP_HD static void madd(xyzz_t &acc, const te_pre_t &q) {
        fp a, b, c, d, e, f, g, h;
        F::template sub_raw<3>(a, acc.y, acc.x);  // Y1 - X1            L 3     V 5
        E::M(a, a, q.ymx);                        // A          3 x 1   L 1     V 2
        F::add_raw(b, acc.y, acc.x);              // Y1 + X1            L 2     V 4
        E::M(b, b, q.ypx);                        // B          2 x 1
        E::M(c, acc.zzz, q.kt);                   // C          1 x 1
        F::add_raw(d, acc.zz, acc.zz);            // D = 2 Z1           L 2     V 4
        F::template sub_raw<3>(e, b, a);          // E = B - A          L 3     V 5
        F::template sub<3>(f, d, c);              // F = D - C          L 1     V 7     (normalised: it meets E and G)
        F::add_raw(g, d, c);                      // G = D + C          L 3     V 6
#if TE_LAZY_NORM >= 2
        F::add_raw(h, b, a);                      // H = B + A          L 2     V 4     (3 x 2 = 6: the largest admissible pair)
#else
        F::add(h, b, a);                          // H = B + A          L 1     V 4     (normalised: it meets E and G)
#endif
        E::M(acc.x, e, f);                        // X3 = E F   3 x 1
        E::M(acc.y, g, h);                        // Y3 = G H   3 x 1 (3 x 2)
        E::M(acc.zzz, e, h);                      // T3 = E H   3 x 1 (3 x 2)
        E::M(acc.zz, f, g);                       // Z3 = F G   1 x 3
    }
#else
    FP_HD static void madd(xyzz_t &acc, const te_pre_t &q) {
        fp a, b, c, d, e, f, g, h;
        F::template sub<3>(a, acc.y, acc.x);      // Y1 - X1                    5
        E::M(a, a, q.ymx);                        // A                          2
        F::add(b, acc.y, acc.x);                  // Y1 + X1                    4
        E::M(b, b, q.ypx);                        // B                          2
        E::M(c, acc.zzz, q.kt);                   // C = T1 * 2 d t2            2
        F::add(d, acc.zz, acc.zz);                // D = 2 Z1                   4
        F::template sub<3>(e, b, a);              // E = B - A                  5
        F::template sub<3>(f, d, c);              // F = D - C                  7
        F::add(g, d, c);                          // G = D + C                  6
        F::add(h, b, a);                          // H = B + A                  4
        E::M(acc.x, e, f);                        // X3 = E F
        E::M(acc.y, g, h);                        // Y3 = G H
        E::M(acc.zzz, e, h);                      // T3 = E H
        E::M(acc.zz, f, g);                       // Z3 = F G
    }
#endif
    // a += b, both extended (add-2008-hwcd-3: 9 M); also doubles (b may equal a)
    FP_HD static void add(xyzz_t &a, const xyzz_t &b) {
        fp A, B, Cc, D, e, f, g, h, t, k;
        F::template sub<3>(A, a.y, a.x);
        F::template sub<3>(t, b.y, b.x);
        E::M(A, A, t);                            // A = (Y1 - X1)(Y2 - X2)     5 x 5
        F::add(B, a.y, a.x);
        F::add(t, b.y, b.x);
        E::M(B, B, t);                            // B = (Y1 + X1)(Y2 + X2)     4 x 4
        load_const(k, 0);
        E::M(Cc, a.zzz, b.zzz);
        E::M(Cc, Cc, k);                          // C = 2 d T1 T2
        E::M(D, a.zz, b.zz);
        F::add(D, D, D);                          // D = 2 Z1 Z2                4
        F::template sub<3>(e, B, A);
        F::template sub<3>(f, D, Cc);
        F::add(g, D, Cc);
        F::add(h, B, A);
        E::M(a.x, e, f);
        E::M(a.y, g, h);
        E::M(a.zzz, e, h);
        E::M(a.zz, f, g);
    }
    FP_HD static void dbl(xyzz_t &a) {
        xyzz_t b = a;
        add(a, b);
    }

    // short-Weierstrass affine (x, y) -> (numerators, denominator) of the map; the caller inverts den (batched):
    //   x' = nu * (dv / den),  y' = nv * (y / den),  den = y * dv
    FP_HD static void map_prepare(fp &nu, fp &nv, fp &dv, fp &den, const fp &x, const fp &y) {
        fp one, t1, xm, cf, cs;
        F::set_one(one);
        load_const(cf, 1);
        load_const(cs, 2);
        F::add(t1, x, one);                       // x + 1                      2
        E::M(nu, t1, cf);                         // f (x + 1)
        E::M(xm, t1, cs);                         // s (x + 1)
        F::add(dv, xm, one);                      // s (x + 1) + 1              3
        F::template sub<2>(nv, xm, one);          // s (x + 1) - 1              4
        E::M(den, y, dv);
    }
    // inv = 1 / den  ->  table entry
    FP_HD static void map_finish(te_entry_t &out, const fp &nu, const fp &nv, const fp &dv, const fp &y, const fp &inv, bool ark = false) {
        fp iy, idv, u, v, k;
        E::M(iy, inv, dv);                        // 1 / y
        E::M(idv, inv, y);                        // 1 / dv
        E::M(u, nu, iy);
        E::M(v, nv, idv);
        te_pre_t q;
        F::template sub<3>(q.ymx, v, u);
        F::add(q.ypx, v, u);
        load_const(k, 0);
        E::M(q.kt, u, v);
        E::M(q.kt, q.kt, k);
        pack(out, q, ark);
    }
    // extended point -> short-Weierstrass affine (canonical); false for the neutral element.  One inversion.
    FP_HD static bool to_sw_affine(fp &x, fp &y, const xyzz_t &a) {
        if (is_inf(a)) return false;
        fp zmy, zpy, den, inv, t, one, c3, cy;
        F::template sub<3>(zmy, a.zz, a.y);       // Z - Y
        F::add(zpy, a.zz, a.y);                   // Z + Y
        E::M(den, a.x, zmy);
        F::inv(inv, den);
        E::M(t, zpy, inv);                        // (Z + Y) / (X (Z - Y))
        load_const(c3, 3);
        load_const(cy, 4);
        F::set_one(one);
        E::M(x, t, a.x);                          // (Z + Y) / (Z - Y)
        E::M(x, x, c3);
        F::template sub<2>(x, x, one);            // sqrt3 xm - 1
        E::M(y, t, a.zz);
        E::M(y, y, cy);
        F::reduce(x);
        F::reduce(y);
        return true;
    }
    // table entry -> extended point (2x' : 2y' : 2 : .)  (tests: read the bases back)
    FP_HD static void entry_to_point(xyzz_t &a, const te_entry_t &e, bool ark = false) {
        te_pre_t q;
        unpack(q, e, ark);
        fp one;
        F::set_one(one);
        F::template sub<2>(a.x, q.ypx, q.ymx);
        F::add(a.y, q.ypx, q.ymx);
        F::add(a.zz, one, one);
        F::set_zero(a.zzz);                       // T is not needed by to_sw_affine
    }
};

// ---------------------------------------------------------------------------------------------
// The accumulate kernel's mixed addition in the 13 x 29-bit field (fp29.cuh): the same add-2008-hwcd-3, every operand
// normalised (value bounds in units of p beside each line; the top limb takes up to 9p).
struct te29_t { fp29 x, y, z, t; };

#if defined(__CUDACC__) && defined(EC_MUL_BY_CALL)
template <class C> __device__ __noinline__ fp29 fp29_mul_call(fp29 a, fp29 b) { fp29 r; Fp29::mul(r, a, b); return r; }
#endif

template <class C> struct Te29T {       // a template only so that its host / device members are instantiated where they are used
    typedef Fp29 F;
#if defined(__CUDACC__) && defined(EC_MUL_BY_CALL)
    __device__ __forceinline__ static void M(fp29 &r, const fp29 &a, const fp29 &b) { r = fp29_mul_call<C>(a, b); }
#else
    FP_HD static void M(fp29 &r, const fp29 &a, const fp29 &b) { F::mul(r, a, b); }
#endif
    FP_HD static void set_inf(te29_t &a) { F::set_zero(a.x); F::set_one(a.y); F::set_one(a.z); F::set_zero(a.t); }
    // ymx, ypx, kt: the three entry values (2^406 form, canonical); negate: -(x, y) = (-x, y)
    FP_HD static void madd(te29_t &acc, const fp29 &ymx, const fp29 &ypx, const fp29 &kt_in, bool negate) {
        fp29 a, b, c, d, e, f, g, h, kt;
        if (negate) F::template neg<2>(kt, kt_in); else kt = kt_in;     //              2
        F::template sub<3>(a, acc.y, acc.x);      // Y1 - X1                    5
        M(a, a, negate ? ypx : ymx);              // A                          1.1
        F::add(b, acc.y, acc.x);                  // Y1 + X1                    2.2
        M(b, b, negate ? ymx : ypx);              // B
        M(c, acc.t, kt);                          // C = T1 * 2 d t2
        F::add(d, acc.z, acc.z);                  // D = 2 Z1                   2.2
        F::template sub<3>(e, b, a);              // E = B - A                  4.1
        F::template sub<3>(f, d, c);              // F = D - C                  5.2
        F::add(g, d, c);                          // G = D + C                  3.3
        F::add(h, b, a);                          // H = B + A                  2.2
        M(acc.x, e, f);                           // X3 = E F
        M(acc.y, g, h);                           // Y3 = G H
        M(acc.t, e, h);                           // T3 = E H
        M(acc.z, f, g);                           // Z3 = F G
    }
    // bucket sum -> the representation of the reduction kernels (14 x 28, 2^392): one multiply per coordinate
    FP_HD static void export_to(xyzz_t &out, const te29_t &a) {
        F::to_fp28(out.x, a.x); F::to_fp28(out.y, a.y); F::to_fp28(out.zz, a.z); F::to_fp28(out.zzz, a.t);
    }
    // the reverse (tests: continue an accumulation held in the cold representation)
    FP_HD static void import_from(te29_t &a, const xyzz_t &in) {
        fp c, v;
        Te<C>::load_const(c, 5);
        const fp *src[4] = {&in.x, &in.y, &in.zz, &in.zzz};
        fp29 *dst[4] = {&a.x, &a.y, &a.z, &a.t};
        for (int k = 0; k < 4; k++) {
            v = *src[k];
            Fp<C>::norm(v);
            Fp<C>::mul(v, v, c);
            Fp<C>::csub_p(v);
            F::repack_from_fp28(*dst[k], v);
        }
    }
};
typedef Te29T<Bls12_377> Te29;

// ---------------------------------------------------------------------------------------------
// The same mixed addition on the FP64 pipe (fpd.cuh: 8 x 48-bit limbs in doubles, R = 2^384): the accumulate kernel's
// inner loop on BLS12-377.  Limb growth in units of 2^48 is given per line: every multiply operand stays below 4 units
// (2^50) and the largest value product is 7 x 6 (limit 152), so the loop needs no normalisation at all.
struct ted_t { fpd x, y, z, t; };
struct ted_pre_t { fpd ymx, ypx, kt; };

#if defined(__CUDACC__) && defined(EC_MUL_BY_CALL)
template <class C> __device__ __noinline__ fpd fpd_mul_call(fpd a, fpd b) { fpd r; FpD<C>::mul(r, a, b); return r; }
#endif

template <class C> struct TeD {
    typedef FpD<C> F;
#if defined(__CUDACC__) && defined(EC_MUL_BY_CALL)
    __device__ __forceinline__ static void M(fpd &r, const fpd &a, const fpd &b) { r = fpd_mul_call<C>(a, b); }
#else
    FP_HD static void M(fpd &r, const fpd &a, const fpd &b) { F::mul(r, a, b); }
#endif
    FP_HD static void set_inf(ted_t &a) { F::set_zero(a.x); F::set_one(a.y); F::set_one(a.z); F::set_zero(a.t); }
    FP_HD static void unpack(ted_pre_t &q, const te_entry_t &e) {
        fp_abi t;
#pragma unroll
        for (int i = 0; i < 12; i++) t.w[i] = e.w[i];
        F::from_abi(q.ymx, t);
#pragma unroll
        for (int i = 0; i < 12; i++) t.w[i] = e.w[12 + i];
        F::from_abi(q.ypx, t);
#pragma unroll
        for (int i = 0; i < 12; i++) t.w[i] = e.w[24 + i];
        F::from_abi(q.kt, t);
    }
    FP_HD static void negate(ted_pre_t &q) {
        fpd t = q.ymx; q.ymx = q.ypx; q.ypx = t;
        F::template neg<2>(q.kt, q.kt);                       // 2 units, value <= 2p
    }
    FP_HD static void madd(ted_t &acc, const ted_pre_t &q) {
        fpd a, b, c, d, e, f, g, h;
        F::template sub<3>(a, acc.y, acc.x);      // Y1 - X1          3 units   5 p
        M(a, a, q.ymx);                           // A                1         2
        F::add(b, acc.y, acc.x);                  // Y1 + X1          2         4
        M(b, b, q.ypx);                           // B                1         2
        M(c, acc.t, q.kt);                        // C                1         2     (q.kt: up to 2 units when negated)
        F::add(d, acc.z, acc.z);                  // D = 2 Z1         2         4
        F::template sub<3>(e, b, a);              // E = B - A        3         5
        F::template sub<3>(f, d, c);              // F = D - C        < 4       7
        F::add(g, d, c);                          // G = D + C        3         6
        F::add(h, b, a);                          // H = B + A        2         4
        M(acc.x, e, f);
        M(acc.y, g, h);
        M(acc.t, e, h);
        M(acc.z, f, g);
    }
    // extended point in the FP64 representation -> the integer representation the reduction kernels use
    // (14 x 28-bit, R = 2^392): through the canonical arkworks form, one multiply each way per coordinate
    FP_HD static void to_int(xyzz_t &out, const ted_t &a) {
        fp_abi t;
        F::to_abi(t, a.x); Fp<C>::from_abi(out.x, t);
        F::to_abi(t, a.y); Fp<C>::from_abi(out.y, t);
        F::to_abi(t, a.z); Fp<C>::from_abi(out.zz, t);
        F::to_abi(t, a.t); Fp<C>::from_abi(out.zzz, t);
    }
};

}  // namespace b200msm

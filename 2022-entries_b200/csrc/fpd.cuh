// fpd.cuh -- base-field arithmetic on the FP64 pipe: 8 limbs x 48 bits held in doubles, Montgomery R = 2^384.
//
// Why (measured on B200, profiles/r01_mul_bench.txt, profiles/r01_fp64_accumulate.txt): the integer multiplier (fp.cuh)
// is bound by the "fmaheavy" half of the FMA pipe -- IMAD.WIDE issues at 32 lanes / clock / SM.  B200 (unlike B300 and
// the consumer parts) has a full-rate FP64 pipe: DFMA at 64 lanes / clock / SM.  A double-precision FMA with a
// rounding-toward-zero bias constant returns either the high or the low 48 bits of a 50 x 50-bit product exactly
// (Emmart, Zheng, Weems, "Faster modular exponentiation using double precision floating point arithmetic on the GPU",
// ARITH 2018):
//      hi = fma_rz(a, b, 2^100)            = 2^100 + floor(a b / 2^48) 2^48      (ulp of [2^100, 2^101) is 2^48)
//      lo = fma_rz(a, b, 2^100 + 2^52 - hi) = 2^52 + (a b mod 2^48)               (exact: the sum fits 53 bits)
// Both results have a FIXED exponent, so their IEEE bit patterns add as 64-bit integers (ALU pipe: IADD3 + IADD3.X)
// and the exponent fields are cancelled by a per-column constant.  A multiply is 128 limb products x (2 DFMA + 1 DADD)
// on the FP64 pipe and ~350 ALU instructions.  MEASURED: 10.7 SM-clocks per multiply against 13.4 for the schoolbook
// integer multiplier and 12.1 for the Karatsuba one -- the 64-bit operand traffic makes it dispatch-bound at 59 % FP64 /
// 48 % ALU / 56 % issue, and inside the accumulate kernel (te.cuh TeD) it ties with the integer flavour instead of
// beating it.  It is therefore a selectable flavour (B200MSM_ACC=fp64), not the default.
// No reference entry does this (all use 32-bit integer multipliers: yrrid MP.cu:141-367, matter-labs
// ff_dispatch_st.cuh:234-481, sppark mont_t.cuh:353-427) -- their target GPUs (A40, sm_70..86) have 1/32-rate FP64.
//
// Conventions
//   value(a) = sum_i a.l[i] 2^(48 i), every limb a non-negative integer-valued double.
//   "Normalised": l[0..6] < 2^48.  mul/sqr accept limbs below 2^50 ("4 units") and value(a) value(b) < 2^384 p
//   (ka kb <= 152 on BLS12-377 for a < ka p, b < kb p); results are normalised with value < 2p.
//   Montgomery form x~ = x 2^384 mod p -- the arkworks form, so the ABI conversion is a bit re-slice, no multiply.
//   add is limb-wise and NOT normalised; sub<K>(a, b) = a + K p - b with K p "limb dominant" (limbs 0..6 >= 2^48 - 1)
//   needs b normalised, value(b) < (K-1) p.  ec.cuh tracks limb growth in units of 2^48 next to the value bounds.
//   BLS12-381 has 3 spare bits in 384 (ka kb <= 4): no room for lazy values, it stays on the integer path (fp.cuh).
//
// The same source compiles for the host (tests/host_emul, with the FPU set to round toward zero); the product never
// runs the host build.
#pragma once
#include "fp.cuh"
#if !defined(__CUDA_ARCH__)
#include <cstring>
#endif

namespace b200msm {

constexpr int FD_LIMBS = 8;
constexpr int FD_BITS = 48;
constexpr uint64_t FD_MASK = (1ull << FD_BITS) - 1;
constexpr uint64_t FD_EXP52 = 0x4330000000000000ull;      // bit pattern of 2^52: mantissa ulp 1
constexpr uint64_t FD_EXP100 = 0x4630000000000000ull;     // bit pattern of 2^100: mantissa ulp 2^48

struct fpd { double l[FD_LIMBS]; };     // 64 bytes

FP_HD double fd_fma_rz(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rz(a, b, c);
#else
    return __builtin_fma(a, b, c);      // host emulation runs with fesetround(FE_TOWARDZERO)
#endif
}
FP_HD uint64_t fd_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
FP_HD double fd_from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}
// exact conversions for integers below 2^52
FP_HD double fd_from_u64(uint64_t v) { return fd_from_bits(v | FD_EXP52) - 0x1p52; }
FP_HD uint64_t fd_to_u64(double d) { return fd_bits(d + 0x1p52) & ((1ull << 52) - 1); }

// 48-bit limb i of the integer whose 28-bit limbs are given by f (constexpr re-slicing of the tables in fp.cuh)
template <class C, int WHICH> FP_HD constexpr uint64_t fd_limb48(int i) {
    uint64_t v = 0;
    for (int b = 0; b < FD_BITS; b++) {
        const int bit = FD_BITS * i + b, li = bit / FP_BITS, lb = bit % FP_BITS;
        if (li < FP_LIMBS) {
            const uint32_t w = WHICH == 0 ? C::p(li) : C::cout(li);      // 0: p, 1: 2^384 mod p (the Montgomery one)
            v |= (uint64_t)((w >> lb) & 1u) << b;
        }
    }
    return v;
}
template <class C> FP_HD constexpr uint64_t fd_p(int i) { return fd_limb48<C, 0>(i); }
template <class C> FP_HD constexpr uint64_t fd_one(int i) { return fd_limb48<C, 1>(i); }
// -p^-1 mod 2^48 and p^-1 mod 2^48 (Newton iteration on the low limb)
template <class C> FP_HD constexpr uint64_t fd_pinv() {
    const uint64_t p0 = fd_p<C>(0);
    uint64_t x = 1;
    for (int k = 0; k < 6; k++) x = x * (2 - p0 * x);
    return x & FD_MASK;
}
template <class C> FP_HD constexpr uint64_t fd_ninv() { return (0 - fd_pinv<C>()) & FD_MASK; }
// limb i of K p in limb-dominant form: limbs 0..6 borrow 2^48 from the limb above
template <class C> FP_HD constexpr uint64_t fd_kp_dom(int k, int i) {
    unsigned __int128 carry = 0;
    uint64_t limb = 0;
    for (int j = 0; j <= i; j++) {
        unsigned __int128 t = (unsigned __int128)fd_p<C>(j) * (uint64_t)k + carry;
        limb = (uint64_t)(t & FD_MASK);
        carry = t >> FD_BITS;
        if (j == FD_LIMBS - 1) limb = (uint64_t)t;
    }
    if (i < FD_LIMBS - 1) limb += (1ull << FD_BITS);
    if (i > 0) limb -= 1;
    return limb;
}
// how many limb products land in column k of an 8 x 8 schoolbook product
FP_HD constexpr int fd_cl(int k) { return (k < 0 || k > 2 * FD_LIMBS - 2) ? 0 : (k < FD_LIMBS ? k + 1 : 2 * FD_LIMBS - 1 - k); }
// start value of column k's accumulator: minus the exponent fields of everything mul() will add into it
// (per product pair (a b, m p): one low half into column k, one high half into column k + 1)
FP_HD constexpr uint64_t fd_col_init(int k) {
    return 0ull - (2ull * (uint64_t)fd_cl(k) * FD_EXP52 + 2ull * (uint64_t)fd_cl(k - 1) * FD_EXP100);
}

template <class C> struct FpD {
    static constexpr int N = FD_LIMBS;
    typedef fpd elem;

    FP_HD static void set_zero(fpd &r) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0.0;
    }
    FP_HD static void set_one(fpd &r) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = (double)fd_one<C>(i);
    }
    FP_HD static bool is_literal_zero(const fpd &a) {
        double o = 0.0;
#pragma unroll
        for (int i = 0; i < N; i++) o += a.l[i];          // limbs are non-negative
        return o == 0.0;
    }
    // carry-propagate so limbs 0..6 are below 2^48 (limbs below 2^52 on entry)
    FP_HD static void norm(fpd &a) {
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < N - 1; i++) {
            const uint64_t v = fd_to_u64(a.l[i]) + c;
            a.l[i] = fd_from_u64(v & FD_MASK);
            c = v >> FD_BITS;
        }
        a.l[N - 1] = fd_from_u64(fd_to_u64(a.l[N - 1]) + c);
    }
    // r = a + b limb-wise (not normalised: limb units add)
    FP_HD static void add(fpd &r, const fpd &a, const fpd &b) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = a.l[i] + b.l[i];
    }
    // r = a + K p - b limb-wise (not normalised: units(a) + 2).  b normalised, value(b) < (K-1) p.
    template <int K> FP_HD static void sub(fpd &r, const fpd &a, const fpd &b) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = a.l[i] + ((double)fd_kp_dom<C>(K, i) - b.l[i]);
    }
    template <int K> FP_HD static void neg(fpd &r, const fpd &a) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = (double)fd_kp_dom<C>(K, i) - a.l[i];
    }

    // one limb product into the column accumulators: lo -> t_lo, hi -> t_hi (bit patterns, see header)
    FP_HD static void prod(uint64_t &t_lo, uint64_t &t_hi, double x, double y) {
        const double hi = fd_fma_rz(x, y, 0x1p100);
        const double lo = fd_fma_rz(x, y, (0x1p100 + 0x1p52) - hi);
        t_lo += fd_bits(lo);
        t_hi += fd_bits(hi);
    }
    // Montgomery product a b / 2^384 mod p.  Row i: columns += a * b_i; m = -t_0 / p mod 2^48; columns += m * p;
    // column 0 is now a multiple of 2^48: its carry moves up and the columns are renamed down by one.
    FP_HD static void mul(fpd &r, const fpd &a, const fpd &b) {
        uint64_t t[N + 1];
#pragma unroll
        for (int j = 0; j <= N; j++) t[j] = fd_col_init(j);
#pragma unroll
        for (int i = 0; i < N; i++) {
            const double bi = b.l[i];
#pragma unroll
            for (int j = 0; j < N; j++) prod(t[j], t[j + 1], a.l[j], bi);
            const uint64_t m = (t[0] * fd_ninv<C>()) & FD_MASK;     // low 48 bits of column 0 are already exact
            const double md = fd_from_u64(m);
#pragma unroll
            for (int j = 0; j < N; j++) prod(t[j], t[j + 1], md, (double)fd_p<C>(j));
            const uint64_t c = t[0] >> FD_BITS;                     // column i is complete: exact, divisible by 2^48
#pragma unroll
            for (int j = 0; j < N; j++) t[j] = t[j + 1];
            t[0] += c;
            t[N] = fd_col_init(i + 1 + N);
        }
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < N - 1; j++) {
            const uint64_t v = t[j] + c;
            r.l[j] = fd_from_u64(v & FD_MASK);
            c = v >> FD_BITS;
        }
        r.l[N - 1] = fd_from_u64(t[N - 1] + c);
    }
    FP_HD static void sqr(fpd &r, const fpd &a) { mul(r, a, a); }

    // a normalised, value < 2p:  if (a >= p) a -= p   (integer domain)
    FP_HD static void csub_p(fpd &a) {
        uint64_t d[N], v[N];
        uint64_t borrow = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            v[i] = fd_to_u64(a.l[i]);
            const uint64_t t = v[i] - fd_p<C>(i) - borrow;
            borrow = t >> 63;
            d[i] = (i < N - 1) ? (t & FD_MASK) : t;
        }
#pragma unroll
        for (int i = 0; i < N; i++) a.l[i] = fd_from_u64(borrow ? v[i] : d[i]);
    }
    // canonical representative in [0, p) of a lazily reduced value (limbs below 2^50)
    FP_HD static void reduce(fpd &a) {
        fpd one, t;
        set_one(one);
        mul(t, a, one);          // a * R / R = a, now < 2p and normalised
        csub_p(t);
        a = t;
    }
    // a == 0 (mod p) for a lazily reduced a (limbs below 2^52, value below 2^7 p).  A multiple k p has low limb
    // k p_0 mod 2^48 with k small: one integer multiply filters all but 2^-32 of the non-zero values.
    FP_HD static bool is_zero(const fpd &a) {
        const uint64_t l0 = fd_to_u64(a.l[0]) & FD_MASK;
        const uint64_t k = (l0 * fd_pinv<C>()) & FD_MASK;
        if (k >= (1u << 16)) return false;
        fpd t = a;
        reduce(t);
        return is_literal_zero(t);
    }

    // arkworks ABI (6 x u64 = 12 x u32, R = 2^384, canonical) <-> 8 x 48-bit limbs: same Montgomery form, bit re-slice
    FP_HD static void from_abi(fpd &r, const fp_abi &a) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            const int bit = FD_BITS * i, w = bit >> 5, s = bit & 31;     // s is 0 or 16
            uint64_t v = (uint64_t)a.w[w] >> s;
            v |= (uint64_t)a.w[w + 1] << (32 - s);
            r.l[i] = fd_from_u64(v & FD_MASK);
        }
    }
    // a lazily reduced -> canonical ABI element
    FP_HD static void to_abi(fp_abi &r, const fpd &a) {
        fpd t = a;
        reduce(t);
        uint64_t v[N];
#pragma unroll
        for (int i = 0; i < N; i++) v[i] = fd_to_u64(t.l[i]);
#pragma unroll
        for (int w = 0; w < 12; w++) {
            const int bit = 32 * w, i = bit / FD_BITS, s = bit % FD_BITS;   // s in {0, 32, 16}
            uint64_t x = v[i] >> s;
            if (s > 16 && i + 1 < N) x |= v[i + 1] << (FD_BITS - s);
            r.w[w] = (uint32_t)x;
        }
    }

    FP_HD static constexpr uint64_t pm2_limb_const(int i) {       // 48-bit limb i of p - 2
        uint64_t borrow = 2, out = 0;
        for (int j = 0; j <= i; j++) {
            const uint64_t pj = fd_p<C>(j);
            out = (pj - borrow) & FD_MASK;
            borrow = (pj < borrow) ? 1u : 0u;
        }
        return out;
    }
    FP_HD static uint64_t pm2_limb(uint32_t w) {
        uint64_t v = 0;
#pragma unroll
        for (int i = 0; i < N; i++)
            if ((uint32_t)i == w) v = pm2_limb_const(i);
        return v;
    }
    // a^(p-2) = 1/a (a normalised, value < 2p; result normalised, < 2p)
    FP_HD static void inv(fpd &r, const fpd &a) {
        fpd acc;
        set_one(acc);
        for (int i = N * FD_BITS - 1; i >= 0; i--) {
            mul(acc, acc, acc);
            if ((pm2_limb(i / FD_BITS) >> (i % FD_BITS)) & 1) mul(acc, acc, a);
        }
        r = acc;
    }
};

}  // namespace b200msm

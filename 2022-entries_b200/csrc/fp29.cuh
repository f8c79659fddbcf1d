// fp29.cuh -- the accumulate kernel's field for BLS12-377: 13 limbs x 29 bits, Montgomery with R = 2^406 (14 rounds).
//
// Why a second representation next to fp.cuh's 14 x 28 (which every other kernel keeps): k_accumulate sits on the practical
// roof of the integer-multiply pipe (fmaheavy 88 % busy against 89 % for the bare multiplier loop, profiles/
// r02_ncu_accumulate_lazy.txt; more warps per scheduler change nothing, profiles/r02_accumulate_variants.txt), so the only way
// down is fewer IMAD.WIDE per multiplication.  p has 377 = 13 x 29 bits: with 29-bit limbs a column of 13 + 13 products
// (each < 2^58) still fits 64 bits, the operand product is 13 x 13 instead of 14 x 14 -- 134 IMAD.WIDE with one level of
// Karatsuba (6 + 7 limbs: 36 + 49 + 49) against 147 -- and a reduction round multiplies 12 limbs of p instead of 13.
// The reduction keeps 14 rounds (R = 2^406 = 2^29 p (roughly)): the twisted-Edwards mixed add feeds the multiplier lazily reduced sums
// (up to 7p x 6p), and the Montgomery bound  a b / R + p  only stays below 2p with that headroom.  302 IMAD.WIDE per
// multiplication against 329.
//
// Conventions (differences from fp.cuh only)
//   value(a) = sum_i a.l[i] 2^(29 i); normalised: l[0..11] < 2^29, the top limb l[12] holds the rest: values up to 9p fit
//   (2^(348 + 32)); the largest the mixed add forms is 7p.
//   mul: BOTH operands normalised (the Karatsuba middle column, 7 products of sums of two limbs, leaves no room for
//   un-normalised limbs: (28 + 7 + 13) 2^58 < 2^64); value(a) value(b) < 2^29 p^2; result normalised, < p + a b / R < 1.1 p.
//   Table entries of the integer accumulate path hold their three values in THIS Montgomery form (2^406), so unpacking an
//   entry stays pure bit slicing; the cold kernels convert (te.cuh Te::unpack form 2) and the bucket sums are converted
//   back to fp.cuh's form when a thread stores its result (Te29::export_to: 4 multiplications per bucket).
// The same source compiles for the host: tests/host_emul checks it against the CPU oracle.
#pragma once
#include "fp.cuh"

namespace b200msm {

constexpr int F29_LIMBS = 13;
constexpr int F29_BITS = 29;
constexpr int F29_ROUNDS = 14;
constexpr uint32_t F29_MASK = (1u << F29_BITS) - 1;

struct fp29 { uint32_t l[F29_LIMBS]; };      // 52 bytes

// BLS12-377 only (p = 1 mod 2^46: p_0 = 1 and -p^-1 = -1 mod 2^29); tests/test_host_emul.py re-derives the tables from p
struct P377x29 {
    FP_HD static constexpr uint32_t p(int i) {
        constexpr uint32_t v[13] = {0x00000001u, 0x08460000u, 0x00000021u, 0x16ba8860u, 0x14800170u, 0x1117dd04u, 0x0e3c7bcdu,
                                    0x1e601ea2u, 0x1b1a22d9u, 0x03650a49u, 0x118ec170u, 0x0f8a21d5u, 0x1ae3a461u};
        return v[i];
    }
    FP_HD static constexpr uint32_t one(int i) {         // 2^406 mod p
        constexpr uint32_t v[13] = {0x19eaf730u, 0x171ffffeu, 0x0d714cf8u, 0x044e31d8u, 0x1eb6e262u, 0x0bfad163u, 0x00d46e9cu,
                                    0x10b6ddbfu, 0x13b9cbddu, 0x075782afu, 0x03bd1d8au, 0x1557cab6u, 0x15742a14u};
        return v[i];
    }
    FP_HD static constexpr uint32_t to392(int i) {       // 2^392 mod p as a plain number: mul(x 2^406, .) = x 2^392
        constexpr uint32_t v[13] = {0x1fff67acu, 0x1907ffffu, 0x1fec35c9u, 0x09d2407fu, 0x05249b0eu, 0x18f0b91au, 0x0aa5575du,
                                    0x16457f13u, 0x16bb2833u, 0x04bcc708u, 0x1a4df717u, 0x02961a87u, 0x03b42e33u};
        return v[i];
    }
};

// limb i of K p with limbs 0..11 each borrowing 2 * 2^29 from the limb above ("limb dominant", as fp.cuh kp_dom; twice the
// borrow because the subtrahend's limbs may exceed 2^29 by a carry: see norm)
FP_HD constexpr uint32_t kp_dom29(int k, int i) {
    uint64_t carry = 0;
    uint32_t limb = 0;
    for (int j = 0; j <= i; j++) {
        uint64_t t = (uint64_t)P377x29::p(j) * (uint32_t)k + carry;
        limb = (uint32_t)(t & F29_MASK);
        carry = t >> F29_BITS;
        if (j == F29_LIMBS - 1) limb = (uint32_t)t;
    }
    if (i < F29_LIMBS - 1) limb += (2u << F29_BITS);
    if (i > 0) limb -= 2;
    return limb;
}

struct Fp29 {
    static constexpr int N = F29_LIMBS;
    typedef P377x29 P;

    FP_HD static void set_zero(fp29 &r) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
    }
    FP_HD static void set_one(fp29 &r) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::one(i);
    }
    // ONE parallel carry step instead of a 12-deep chain: every limb hands its excess over 29 bits to the next one at the
    // same time.  Inputs have limbs below 2^32 - 8 (sums of two operands, or operand + K p - operand), so a carry is at most
    // 7 and the result has limbs 0..11 below 2^29 + 8: "normalised" in this file means that, not < 2^29 -- the multiplier's
    // column bound has the room (48 of 64 units used), and sub / neg borrow twice 2^29 per limb so that such a limb can be
    // subtracted.  Depth 3 instead of 36 dependent instructions: the accumulate kernel spends its non-multiplier time in
    // these chains while the multiply pipe waits (profiles/r02_accumulate_variants.txt).
    FP_HD static void norm(fp29 &a) {
#if defined(F29_SERIAL_NORM)
        uint32_t c = 0;
#pragma unroll
        for (int i = 0; i < N - 1; i++) {
            uint32_t t = a.l[i] + c;
            a.l[i] = t & F29_MASK;
            c = t >> F29_BITS;
        }
        a.l[N - 1] += c;
#else
        uint32_t c[N - 1];
#pragma unroll
        for (int i = 0; i < N - 1; i++) { c[i] = a.l[i] >> F29_BITS; a.l[i] &= F29_MASK; }
#pragma unroll
        for (int i = 1; i < N; i++) a.l[i] += c[i - 1];
#endif
    }
    FP_HD static void add(fp29 &r, const fp29 &a, const fp29 &b) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = a.l[i] + b.l[i];
        norm(r);
    }
    // r = a + K p - b, normalised.  b normalised with value(b) < (K - 1) p; limbs of a below 2^30
    template <int K> FP_HD static void sub(fp29 &r, const fp29 &a, const fp29 &b) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = a.l[i] + kp_dom29(K, i) - b.l[i];
        norm(r);
    }
    template <int K> FP_HD static void neg(fp29 &r, const fp29 &a) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = kp_dom29(K, i) - a.l[i];
        norm(r);
    }

    // Montgomery product a b / 2^406 mod p.  a = a0 + a1 X, X = 2^(6*29): a0 = limbs 0..5, a1 = limbs 6..12 (7 limbs, the top
    // limb alone in its half-sum, so the only full 32 x 32-bit product sits in a column of its own).
    //   z0 = a0 b0 (36), z2 = a1 b1 (49), z1 = (a0 + a1)(b0 + b1) - z0 - z2 (49), then 14 rounds x 12 limbs of p (p_0 = 1).
    FP_HD static void mul(fp29 &r, const fp29 &a, const fp29 &b) {
        constexpr int L = 6, H = 7, T = N + F29_ROUNDS;      // 27 columns
        uint64_t t[T];
        uint64_t mid[2 * H - 1];
        uint32_t sa[H], sb[H];
#pragma unroll
        for (int i = 0; i < L; i++) { sa[i] = a.l[i] + a.l[L + i]; sb[i] = b.l[i] + b.l[L + i]; }
        sa[H - 1] = a.l[N - 1];
        sb[H - 1] = b.l[N - 1];
#pragma unroll
        for (int k = 0; k < T; k++) t[k] = 0;
#if defined(F29_K2)
        // second Karatsuba level on z0 (6 = 3 + 3: 27 products instead of 36) and z2 (7 = 3 + 4, the top limb alone again:
        // 41 instead of 49); the middle product stays schoolbook (its operands are already sums: a third sum would overflow)
        {
            uint32_t ua[3], ub[3];
            uint64_t pm[5];
#pragma unroll
            for (int i = 0; i < 3; i++) { ua[i] = a.l[i] + a.l[3 + i]; ub[i] = b.l[i] + b.l[3 + i]; }
#pragma unroll
            for (int i = 0; i < 3; i++) {
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    if (i == 0 || j == 2) { t[i + j] = (uint64_t)a.l[i] * b.l[j]; t[6 + i + j] = (uint64_t)a.l[3 + i] * b.l[3 + j]; pm[i + j] = (uint64_t)ua[i] * ub[j]; }
                    else { t[i + j] += (uint64_t)a.l[i] * b.l[j]; t[6 + i + j] += (uint64_t)a.l[3 + i] * b.l[3 + j]; pm[i + j] += (uint64_t)ua[i] * ub[j]; }
                }
            }
#pragma unroll
            for (int k = 0; k < 5; k++) pm[k] -= t[k] + t[6 + k];
#pragma unroll
            for (int k = 0; k < 5; k++) t[3 + k] += pm[k];
        }
        {
            uint32_t ua[4], ub[4];
            uint64_t q0[5], qm[7];
            uint64_t *u = t + 2 * L;                        // z2 columns 0..12 live in t[12..24]
#pragma unroll
            for (int i = 0; i < 3; i++) { ua[i] = a.l[L + i] + a.l[L + 3 + i]; ub[i] = b.l[L + i] + b.l[L + 3 + i]; }
            ua[3] = a.l[N - 1];
            ub[3] = b.l[N - 1];
#pragma unroll
            for (int i = 0; i < 4; i++) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (i < 3 && j < 3) {
                        if (i == 0 || j == 2) q0[i + j] = (uint64_t)a.l[L + i] * b.l[L + j];
                        else q0[i + j] += (uint64_t)a.l[L + i] * b.l[L + j];
                    }
                    if (i == 0 || j == 3) { u[6 + i + j] = (uint64_t)a.l[L + 3 + i] * b.l[L + 3 + j]; qm[i + j] = (uint64_t)ua[i] * ub[j]; }
                    else { u[6 + i + j] += (uint64_t)a.l[L + 3 + i] * b.l[L + 3 + j]; qm[i + j] += (uint64_t)ua[i] * ub[j]; }
                }
            }
#pragma unroll
            for (int k = 0; k < 7; k++) qm[k] -= (k < 5 ? q0[k] : 0) + u[6 + k];
#pragma unroll
            for (int k = 0; k < 5; k++) u[k] = q0[k];
            u[5] = 0;
#pragma unroll
            for (int k = 0; k < 7; k++) u[3 + k] += qm[k];
        }
#pragma unroll
        for (int i = 0; i < H; i++) {
#pragma unroll
            for (int j = 0; j < H; j++) {
                if (i == 0 || j == H - 1) mid[i + j] = (uint64_t)sa[i] * sb[j];
                else mid[i + j] += (uint64_t)sa[i] * sb[j];
            }
        }
#else
#pragma unroll
        for (int i = 0; i < H; i++) {
#pragma unroll
            for (int j = 0; j < H; j++) {
                if (i < L && j < L) {
                    if (i == 0 || j == L - 1) t[i + j] = (uint64_t)a.l[i] * b.l[j];
                    else t[i + j] += (uint64_t)a.l[i] * b.l[j];
                }
                if (i == 0 || j == H - 1) {
                    t[2 * L + i + j] = (uint64_t)a.l[L + i] * b.l[L + j];
                    mid[i + j] = (uint64_t)sa[i] * sb[j];
                } else {
                    t[2 * L + i + j] += (uint64_t)a.l[L + i] * b.l[L + j];
                    mid[i + j] += (uint64_t)sa[i] * sb[j];
                }
            }
        }
#endif
        // z1 = mid - z0 - z2 per column (z0 has columns 0..10, z2 and mid 0..12), added at column offset 6
#pragma unroll
        for (int k = 0; k < 2 * H - 1; k++) mid[k] -= (k < 2 * L - 1 ? t[k] : 0) + t[2 * L + k];
#pragma unroll
        for (int k = 0; k < 2 * H - 1; k++) t[L + k] += mid[k];
#pragma unroll
        for (int i = 0; i < F29_ROUNDS; i++) {
            const uint32_t m = opaque32((0u - (uint32_t)t[i]) & F29_MASK);      // -p^-1 = -1 mod 2^29
#pragma unroll
            for (int j = 0; j < N; j++) t[i + j] += (uint64_t)m * P::p(j);
            t[i + 1] += t[i] >> F29_BITS;
        }
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < N - 1; j++) {
            uint64_t v = t[F29_ROUNDS + j] + c;
            r.l[j] = (uint32_t)v & F29_MASK;
            c = v >> F29_BITS;
        }
        r.l[N - 1] = (uint32_t)(t[T - 1] + c);
    }
    // a normalised, value < 2p: if (a >= p) a -= p
    FP_HD static void csub_p(fp29 &a) {
        uint32_t d[N];
        uint32_t borrow = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            uint32_t t = a.l[i] - P::p(i) - borrow;
            borrow = t >> 31;
            d[i] = (i < N - 1) ? (t & F29_MASK) : t;
        }
#pragma unroll
        for (int i = 0; i < N; i++) a.l[i] = borrow ? a.l[i] : d[i];
    }
    // 12 x 32-bit packed canonical value <-> 13 x 29-bit limbs: bit slicing only
    FP_HD static void repack_from_abi(fp29 &r, const fp_abi &a) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            const int bit = F29_BITS * i, w = bit >> 5, s = bit & 31;
            uint64_t v = a.w[w];
            if (w + 1 < 12) v |= (uint64_t)a.w[w + 1] << 32;
            r.l[i] = (uint32_t)(v >> s) & F29_MASK;
        }
    }
    // canonical 13 x 29 value -> fp.cuh's 14 x 28 limbs (bit slicing)
    FP_HD static void repack_to_fp28(fp &r, const fp29 &a) {
#pragma unroll
        for (int i = 0; i < FP_LIMBS; i++) {
            const int bit = FP_BITS * i, k = bit / F29_BITS, s = bit % F29_BITS;
            uint64_t v = (uint64_t)a.l[k] >> s;
            if (k + 1 < N) v |= (uint64_t)a.l[k + 1] << (F29_BITS - s);
            r.l[i] = (uint32_t)v & FP_MASK;
        }
    }
    FP_HD static void repack_from_fp28(fp29 &r, const fp &a) {       // canonical 14 x 28 value -> 13 x 29 limbs
#pragma unroll
        for (int i = 0; i < N; i++) {
            const int bit = F29_BITS * i, k = bit / FP_BITS, s = bit % FP_BITS;
            uint64_t v = (uint64_t)a.l[k] >> s;
            if (k + 1 < FP_LIMBS) v |= (uint64_t)a.l[k + 1] << (FP_BITS - s);
            if (k + 2 < FP_LIMBS && 2 * FP_BITS - s < F29_BITS) v |= (uint64_t)a.l[k + 2] << (2 * FP_BITS - s);
            r.l[i] = (uint32_t)v & F29_MASK;
        }
    }
    // x 2^406 (any representative below 2p... 9p, normalised) -> x 2^392 in fp.cuh's limbs, canonical
    FP_HD static void to_fp28(fp &r, const fp29 &a) {
        fp29 c, t;
#pragma unroll
        for (int i = 0; i < N; i++) c.l[i] = P::to392(i);
        mul(t, a, c);
        csub_p(t);
        repack_to_fp28(r, t);
    }
};

}  // namespace b200msm

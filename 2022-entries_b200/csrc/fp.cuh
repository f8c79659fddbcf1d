// fp.cuh -- 384-bit Montgomery field arithmetic for sm_100a (12 x 32-bit limbs, lazily reduced).
//
// Replaces, for this path, the device field libraries of the reference entries
// (yrrid MP.cu:141-367 + Curve.cu:21-449, matter-labs ff_dispatch_st.cuh:234-481, sppark
// mont_t.cuh:39-58,353-427).  Written from scratch: one even/odd column Montgomery multiply whose
// mad.lo.cc/madc.hi.cc pairs ptxas fuses into IMAD.WIDE.U32(.X) with a predicate carry (checked with
// cuobjdump: 288 IMAD.WIDE per multiply), no final subtraction (the spare bits of the 384-bit
// container absorb it), and subtraction as a + k*p - b with compile-time k*p immediates.
//
// The same source compiles for the host (CARRY flag emulated) so tests/host_emul can check the limb
// logic against the CPU oracle without a GPU.  The product never runs the host build.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FP_HD __host__ __device__ __forceinline__
#define FP_D __device__ __forceinline__
#else
#define FP_HD inline
#define FP_D inline
#endif

namespace b200msm {

// ---------------------------------------------------------------------------------------------
// carry-chain primitives
#if defined(__CUDA_ARCH__)
FP_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FP_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FP_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FP_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FP_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FP_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FP_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
FP_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
FP_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
FP_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
// host emulation of the PTX condition-code register (tests only)
static thread_local uint32_t g_cc = 0;
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + g_cc; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + g_cc; return (uint32_t)t; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; g_cc = (uint32_t)(t >> 32) & 1; return (uint32_t)t; }   // g_cc = borrow
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - g_cc; g_cc = (uint32_t)(t >> 32) & 1; return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - g_cc; return (uint32_t)t; }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)((uint64_t)a * b) + c; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)((uint64_t)a * b) + c + g_cc; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + g_cc; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + g_cc; return (uint32_t)t; }
#endif

// ---------------------------------------------------------------------------------------------
// curve parameters (cross-checked there against sppark bls12-377.hpp:10-39,
// bls12-381.hpp:10-37, ff_config.cuh:13-40, SHM.cu:13-48).  Limb accessors are constexpr functions so
// fully unrolled code sees immediates.
struct Bls12_377 {
    static constexpr int ID = 0;
    static constexpr int SCALAR_BITS = 253;
    static constexpr uint32_t INV = 0xffffffffu;     // -p^-1 mod 2^32
    static constexpr uint32_t P0INV = 0x00000001u;   //  p^-1 mod 2^32
    static constexpr int SLACK = 152;                // floor(2^384 / p): mul(a<ka*p, b<kb*p) < 2p while ka*kb <= SLACK
    static constexpr bool TIGHT = false;             // lazy bounds X<8p, Y<4p, ZZ,ZZZ<2p hold without extra reductions
    static constexpr uint32_t B3 = 3;                // 3*b
    FP_HD static constexpr uint32_t p(int i) {
        constexpr uint32_t v[12] = {0x00000001u, 0x8508c000u, 0x30000000u, 0x170b5d44u, 0xba094800u, 0x1ef3622fu,
                                    0x00f5138fu, 0x1a22d9f3u, 0x6ca1493bu, 0xc63b05c0u, 0x17c510eau, 0x01ae3a46u};
        return v[i];
    }
    FP_HD static constexpr uint32_t one(int i) {     // R mod p
        constexpr uint32_t v[12] = {0xffffff68u, 0x02cdffffu, 0x7fffffb1u, 0x51409f83u, 0x8a7d3ff2u, 0x9f7db3a9u,
                                    0x6e7c6305u, 0x7b4e97b7u, 0x803c84e8u, 0x4cf495bfu, 0xe2fdf49au, 0x008d6661u};
        return v[i];
    }
    FP_HD static constexpr uint32_t r2(int i) {      // R^2 mod p
        constexpr uint32_t v[12] = {0x9400cd22u, 0xb786686cu, 0xb00431b1u, 0x0329fcaau, 0x62d6b46du, 0x22a5f111u,
                                    0x827dc3acu, 0xbfdf7d03u, 0x41790bf9u, 0x837e92f0u, 0x1e914b88u, 0x006dfccbu};
        return v[i];
    }
    FP_HD static constexpr uint32_t r(int i) {       // scalar-field modulus
        constexpr uint32_t v[8] = {0x00000001u, 0x0a118000u, 0xd0000001u, 0x59aa76feu, 0x5c37b001u, 0x60b44d1eu, 0x9a2ca556u, 0x12ab655eu};
        return v[i];
    }
};

struct Bls12_381 {
    static constexpr int ID = 1;
    static constexpr int SCALAR_BITS = 255;
    static constexpr uint32_t INV = 0xfffcfffdu;
    static constexpr uint32_t P0INV = 0x00030003u;
    static constexpr int SLACK = 9;
    static constexpr bool TIGHT = true;              // only 3 spare bits: every add/sub result is brought back below 2p
    static constexpr uint32_t B3 = 12;
    FP_HD static constexpr uint32_t p(int i) {
        constexpr uint32_t v[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                    0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
        return v[i];
    }
    FP_HD static constexpr uint32_t one(int i) {
        constexpr uint32_t v[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                                    0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
        return v[i];
    }
    FP_HD static constexpr uint32_t r2(int i) {
        constexpr uint32_t v[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                                    0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
        return v[i];
    }
    FP_HD static constexpr uint32_t r(int i) {
        constexpr uint32_t v[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return v[i];
    }
};

// limb i of k*p, k small, as a compile-time constant
template <class C> FP_HD constexpr uint32_t kp_limb(int k, int i) {
    uint64_t carry = 0;
    uint32_t out = 0;
    for (int j = 0; j <= i; j++) {
        uint64_t t = (uint64_t)C::p(j) * (uint32_t)k + carry;
        out = (uint32_t)t;
        carry = t >> 32;
    }
    return out;
}

struct fp { uint32_t l[12]; };

template <class C> struct Fp {
    static constexpr int N = 12;

    FP_HD static void set_zero(fp &r) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
    }
    FP_HD static void set_one(fp &r) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = C::one(i);
    }
    FP_HD static bool is_literal_zero(const fp &a) {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; i++) o |= a.l[i];
        return o == 0;
    }

    // r = a + b (no reduction; caller owns the bound)
    FP_HD static void add(fp &r, const fp &a, const fp &b) {
        r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
        r.l[N - 1] = addc(a.l[N - 1], b.l[N - 1]);
    }
    // r = a + K*p - b   (requires b <= K*p)
    template <int K> FP_HD static void sub(fp &r, const fp &a, const fp &b) {
        fp t;
        t.l[0] = add_cc(a.l[0], kp_limb<C>(K, 0));
#pragma unroll
        for (int i = 1; i < N - 1; i++) t.l[i] = addc_cc(a.l[i], kp_limb<C>(K, i));
        t.l[N - 1] = addc(a.l[N - 1], kp_limb<C>(K, N - 1));
        r.l[0] = sub_cc(t.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = subc_cc(t.l[i], b.l[i]);
        r.l[N - 1] = subc(t.l[N - 1], b.l[N - 1]);
    }
    // r = K*p - a  (requires a <= K*p)
    template <int K> FP_HD static void neg(fp &r, const fp &a) {
        r.l[0] = sub_cc(kp_limb<C>(K, 0), a.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = subc_cc(kp_limb<C>(K, i), a.l[i]);
        r.l[N - 1] = subc(kp_limb<C>(K, N - 1), a.l[N - 1]);
    }
    // if (a >= K*p) a -= K*p
    template <int K> FP_HD static void csub(fp &a) {
        fp t;
        t.l[0] = sub_cc(a.l[0], kp_limb<C>(K, 0));
#pragma unroll
        for (int i = 1; i < N; i++) t.l[i] = subc_cc(a.l[i], kp_limb<C>(K, i));
        uint32_t borrow = subc(0, 0);   // 0xffffffff when a < K*p
#pragma unroll
        for (int i = 0; i < N; i++) a.l[i] = borrow ? a.l[i] : t.l[i];
    }
    // canonical representative in [0, p) of a value below 16p
    FP_HD static void reduce(fp &a) {
        if constexpr (!C::TIGHT) { csub<8>(a); csub<4>(a); }
        csub<2>(a);
        csub<1>(a);
    }
    // a == 0 (mod p) for a < 16p.  A multiple k*p has low limb k*p0, so k = a0 * p0^-1 must be small: one
    // multiply filters out all but ~2^-28 of the non-zero values before the exact comparison.
    FP_HD static bool is_zero(const fp &a) {
        uint32_t k = a.l[0] * C::P0INV;
        if (k >= 16) return false;
        fp t = a;
        reduce(t);
        return is_literal_zero(t);
    }

    // Montgomery product a*b/2^384 mod p, lazily reduced:  result < (ka*kb/SLACK + 1) * p for a < ka*p, b < kb*p,
    // and a + p must stay below 2^384.
    //
    // Even/odd column form.  T = ev + od*2^32 with ev at limb positions 0..12 and od[j] at position j+1.  Round i
    // adds a*b_i and m*p (m chosen so limb 0 cancels) with four 6-pair carry chains -- every pair is one
    // IMAD.WIDE.U32.X -- then divides by 2^32 by renaming: od becomes the new ev, ev[2..] the new od, and the one
    // limb that lands on an occupied position (ev[1] onto od[0]) is added with its carry parked in CC until the
    // first madc of the next round.  T never exceeds a + p, so od needs no top carry limb.
    FP_HD static void mul(fp &r, const fp &a, const fp &b) {
        uint32_t ev[N + 1], od[N];
#pragma unroll
        for (int i = 0; i < N; i++) { ev[i] = 0; od[i] = 0; }
        ev[N] = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            const uint32_t bi = b.l[i];
            // od += (a1, a3, ...) * bi, taking in the carry parked by the previous round's merge
            od[0] = (i == 0) ? mad_lo_cc(a.l[1], bi, od[0]) : madc_lo_cc(a.l[1], bi, od[0]);
            od[1] = madc_hi_cc(a.l[1], bi, od[1]);
#pragma unroll
            for (int j = 2; j < N; j += 2) {
                od[j] = madc_lo_cc(a.l[j + 1], bi, od[j]);
                od[j + 1] = (j + 2 < N) ? madc_hi_cc(a.l[j + 1], bi, od[j + 1]) : madc_hi(a.l[j + 1], bi, od[j + 1]);
            }
            // ev += (a0, a2, ...) * bi
            ev[0] = mad_lo_cc(a.l[0], bi, ev[0]);
            ev[1] = madc_hi_cc(a.l[0], bi, ev[1]);
#pragma unroll
            for (int j = 2; j < N; j += 2) {
                ev[j] = madc_lo_cc(a.l[j], bi, ev[j]);
                ev[j + 1] = madc_hi_cc(a.l[j], bi, ev[j + 1]);
            }
            ev[N] = addc(ev[N], 0);
            const uint32_t m = ev[0] * C::INV;
            // ev += (p0, p2, ...) * m   -> ev[0] == 0
            ev[0] = mad_lo_cc(C::p(0), m, ev[0]);
            ev[1] = madc_hi_cc(C::p(0), m, ev[1]);
#pragma unroll
            for (int j = 2; j < N; j += 2) {
                ev[j] = madc_lo_cc(C::p(j), m, ev[j]);
                ev[j + 1] = madc_hi_cc(C::p(j), m, ev[j + 1]);
            }
            ev[N] = addc(ev[N], 0);
            // od += (p1, p3, ...) * m
            od[0] = mad_lo_cc(C::p(1), m, od[0]);
            od[1] = madc_hi_cc(C::p(1), m, od[1]);
#pragma unroll
            for (int j = 2; j < N; j += 2) {
                od[j] = madc_lo_cc(C::p(j + 1), m, od[j]);
                od[j + 1] = (j + 2 < N) ? madc_hi_cc(C::p(j + 1), m, od[j + 1]) : madc_hi(C::p(j + 1), m, od[j + 1]);
            }
            // divide by 2^32: new ev = od (+ ev[1] on limb 0, carry parked), new od = ev >> 64
            const uint32_t e0 = add_cc(od[0], ev[1]);
            uint32_t nod[N];
#pragma unroll
            for (int j = 0; j < N - 1; j++) nod[j] = ev[j + 2];
            nod[N - 1] = 0;
            ev[0] = e0;
#pragma unroll
            for (int j = 1; j < N; j++) ev[j] = od[j];
            ev[N] = 0;
#pragma unroll
            for (int j = 0; j < N; j++) od[j] = nod[j];
        }
        // r = ev + od*2^32, consuming the parked carry
        r.l[0] = ev[0];
#pragma unroll
        for (int j = 1; j < N - 1; j++) r.l[j] = addc_cc(ev[j], od[j - 1]);
        r.l[N - 1] = addc(ev[N - 1], od[N - 2]);
    }
    FP_HD static void sqr(fp &r, const fp &a) { mul(r, a, a); }

    // Montgomery form <-> canonical integer
    FP_HD static void to_mont(fp &r, const fp &a) {
        fp t;
#pragma unroll
        for (int i = 0; i < N; i++) t.l[i] = C::r2(i);
        mul(r, a, t);
        reduce(r);
    }
    FP_HD static void from_mont(fp &r, const fp &a) {
        fp t;
        set_zero(t);
        t.l[0] = 1;
        mul(r, a, t);
        reduce(r);
    }
    // a^(p-2): 1/a for a != 0 (mod p); input below 2p, output below 2p.  Fixed square-and-multiply ladder.
    FP_HD static void inv(fp &r, const fp &a) {
        fp acc;
        set_one(acc);
        // exponent p - 2, scanned from the top bit
        for (int i = 383; i >= 0; i--) {
            sqr(acc, acc);
            uint32_t w = (i >> 5);
            uint32_t e = pm2_limb(w);
            if ((e >> (i & 31)) & 1) mul(acc, acc, a);
        }
        r = acc;
    }
    FP_HD static constexpr uint32_t pm2_const(int i) {   // limb i of p - 2
        uint32_t borrow = 2, out = 0;
        for (int j = 0; j <= i; j++) {
            uint32_t pj = C::p(j);
            out = pj - borrow;
            borrow = (pj < borrow) ? 1u : 0u;
        }
        return out;
    }
    FP_HD static uint32_t pm2_limb(uint32_t w) {
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < N; i++)
            if ((uint32_t)i == w) v = pm2_const(i);
        return v;
    }
    // lazy/tight policy helpers: in lazy mode (BLS12-377) r = a + KL*p - b, bound a + KL*p; in tight mode
    // (BLS12-381) operands are below 2p and the result is brought back below 2p.
    template <int KL> FP_HD static void subx(fp &r, const fp &a, const fp &b) {
        if constexpr (C::TIGHT) { sub<2>(r, a, b); csub<2>(r); } else { sub<KL>(r, a, b); }
    }
    FP_HD static void addx(fp &r, const fp &a, const fp &b) {
        add(r, a, b);
        if constexpr (C::TIGHT) csub<2>(r);
    }
};

}  // namespace b200msm

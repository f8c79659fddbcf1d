// fp.cuh -- base-field arithmetic for sm_100a: 14 limbs x 28 bits ("reduced radix"), Montgomery with R = 2^392.
//
// Why this representation (measured on B200: profiles/r01_mul_bench.txt, profiles/r01_ncu_accumulate_*.txt): integer
// multiplies issue only on the "fmaheavy" half of the FMA pipe, at 32 lanes (one warp instruction) per clock per SM,
// and that pipe is what bounds the whole MSM.  A 12 x 32-bit multiplier (every reference entry) needs carry-predicated
// IMAD.WIDE.X chains plus IMAD.HI/IMAD.X pairs for the constant-operand half -- 576 heavy-pipe slots per multiply.
// With 28-bit limbs a column of 14 + 14 products (each < 2^56) fits a 64-bit accumulator, so the schoolbook multiplier
// is 14*14*2 - 14 = 378 plain IMAD.WIDE with NO carry flags, and one level of Karatsuba on a * b brings it to 329
// (mul_karatsuba below); the carry work left (one shift+add per round, one normalisation at the end, the Karatsuba
// column sums) goes to the ALU pipe, which is a quarter busy.  Stands in for yrrid MP.cu:141-367, matter-labs
// ff_dispatch_st.cuh:234-481, sppark mont_t.cuh:353-427 (all 12 x 32-bit carry-chain multipliers).
//
// Conventions
//   value(a) = sum_i a.l[i] * 2^(28 i).  "Normalised": l[0..12] < 2^28 (l[13] holds whatever is left).
//   Montgomery form x~ = x * 2^392 mod p.  The arkworks ABI uses 6 x u64 with R = 2^384; abi_to_fp / fp_to_abi convert
//   (bit repack + one multiply by 2^400 resp. 2^384), once per base at table build and once per result.
//   mul(a, b): NORMALISED operands (mul_karatsuba below; mul_schoolbook also takes limbs up to 2^30), and
//   value(a) * value(b) < 2^392 * p  (i.e. ka * kb < 2^15 for BLS12-377, 2^11 for BLS12-381 when a < ka p, b < kb p);
//   result normalised, value < 2p.  The 11-15 spare bits make every bound in ec.cuh hold without value reductions.
//   sub<K>(a, b) = a + K p - b limb-wise, K p stored "limb-dominant" (every limb >= 2^28 - 1) so no limb goes
//   negative: requires b normalised and value(b) < (K - 1) p.  add/sub results are NOT normalised; norm() does it.
//
// The same source compiles for the host (plain C++), so tests/host_emul checks it against the CPU oracle without a
// GPU.  The product never runs the host build.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FP_HD __host__ __device__ __forceinline__
#else
#define FP_HD inline
#endif

namespace b200msm {

constexpr int FP_LIMBS = 14;
constexpr int FP_BITS = 28;
constexpr uint32_t FP_MASK = (1u << FP_BITS) - 1;

struct fp { uint32_t l[FP_LIMBS]; };        // 56 bytes
struct fp_abi { uint32_t w[12]; };          // arkworks Fq: 6 x u64 little-endian, Montgomery R = 2^384, canonical

// ---------------------------------------------------------------------------------------------
// curve parameters in 28-bit limbs (SURVEY.md Appendix B values, re-sliced; tests/test_host_emul.py re-derives every
// table below from p with Python integers).  constexpr accessors so unrolled code sees immediates.
struct Bls12_377 {
    static constexpr int ID = 0;
    static constexpr int SCALAR_BITS = 253;          // bits of r
    static constexpr int WINDOW_BITS = 253;          // bits the window plan covers: the scalar is first reduced mod r and its top bit
    static constexpr bool FOLD_SCALAR = true;        // folded away (s -> r - s, point negated) -- both valid only for points of order r
    static constexpr uint32_t INV = 0xfffffffu;      // -p^-1 mod 2^28
    static constexpr uint32_t P0INV = 0x0000001u;    //  p^-1 mod 2^28
    FP_HD static constexpr uint32_t p(int i) {
        constexpr uint32_t v[14] = {0x0000001u, 0x08c0000u, 0x0000085u, 0x5d44300u, 0x800170bu, 0x2fba094u, 0xf1ef362u,
                                    0x00f5138u, 0xa22d9f3u, 0xa1493b1u, 0xb05c06cu, 0x10eac63u, 0xa4617c5u, 0x0001ae3u};
        return v[i];
    }
    FP_HD static constexpr uint32_t one(int i) {     // 2^392 mod p
        constexpr uint32_t v[14] = {0xfff67acu, 0x20fffffu, 0xfb0d727u, 0xe9203ffu, 0x249b0e4u, 0xe172345u, 0x955d771u,
                                    0x2bf89aau, 0xb2833b2u, 0x98e116bu, 0x7dc5c97u, 0x0d43e93u, 0x2e3314bu, 0x00003b4u};
        return v[i];
    }
    FP_HD static constexpr uint32_t cin(int i) {     // 2^400 mod p: mul(repack(x * 2^384), cin) = x * 2^392
        constexpr uint32_t v[14] = {0xf67abddu, 0xcdbffffu, 0x0d714f0u, 0x5fed6fbu, 0x1adbe5bu, 0xebc50d7u, 0x49c2b74u,
                                    0xd718fccu, 0x56fe5f2u, 0xd413a69u, 0xa9348beu, 0xf3d01dcu, 0xb9c0b1bu, 0x000070eu};
        return v[i];
    }
    FP_HD static constexpr uint32_t cout(int i) {    // 2^384 mod p: mul(x * 2^392, cout) = x * 2^384
        constexpr uint32_t v[14] = {0xfffff68u, 0xcdfffffu, 0xfffb102u, 0x9f837ffu, 0xff25140u, 0xa98a7d3u, 0x59f7db3u,
                                    0x6e7c630u, 0xb4e97b7u, 0x3c84e87u, 0x495bf80u, 0xf49a4cfu, 0x661e2fdu, 0x00008d6u};
        return v[i];
    }
    FP_HD static constexpr uint32_t r(int i) {       // scalar-field modulus, 8 x 32-bit
        constexpr uint32_t v[8] = {0x00000001u, 0x0a118000u, 0xd0000001u, 0x59aa76feu, 0x5c37b001u, 0x60b44d1eu, 0x9a2ca556u, 0x12ab655eu};
        return v[i];
    }
    // Twisted-Edwards model of G1 with a = -1 (te.cuh): -x^2 + y^2 = 1 + d x^2 y^2.  All in Montgomery form (2^392);
    // tests/test_host_emul.py re-derives them from p (s = 1/sqrt(3), Montgomery A = -3s, B = s, a1 = (A+2)/B,
    // d1 = (A-2)/B, f = sqrt(-a1), d = -d1/a1 -- the parameters arkworks lists for this curve,
    // P1B/nickray/driver/bls12-377/src/curves/g1.rs:52-113).
    static constexpr bool HAS_TE = true;
    FP_HD static constexpr uint32_t te_k(int i) {    // 2 d
        constexpr uint32_t v[14] = {0x706c521u, 0xdf4444au, 0x8fa7e68u, 0x371b2aau, 0xdd6cd3du, 0xebe579eu, 0x3577a6au, 0x1d3f1c6u, 0x2ada318u, 0x2a73d9du, 0x7da6ef5u, 0xce09495u, 0xfed4ab8u, 0x0001813u};
        return v[i];
    }
    FP_HD static constexpr uint32_t te_f(int i) {    // sqrt(-a1): x scaling that makes a = -1
        constexpr uint32_t v[14] = {0x605e344u, 0x9db8eadu, 0xc7a99c9u, 0x53cf352u, 0x6254f3bu, 0x07ad145u, 0x769ac1au, 0xeb261abu, 0xa5ececdu, 0x625c845u, 0x568b26eu, 0x37d10bau, 0x84097bau, 0x0000867u};
        return v[i];
    }
    FP_HD static constexpr uint32_t te_s(int i) {    // 1 / sqrt(3)
        constexpr uint32_t v[14] = {0x1a0a113u, 0xb0dad83u, 0xc8df2c0u, 0xc9ee451u, 0x1935a6eu, 0x6425e1bu, 0xacf94dfu, 0x9d801ddu, 0x14590dbu, 0xad64a0cu, 0x2a17c35u, 0x6b6b819u, 0x0112a8cu, 0x0000fa6u};
        return v[i];
    }
    FP_HD static constexpr uint32_t te_s3(int i) {   // sqrt(3) = 1 / s
        constexpr uint32_t v[14] = {0x4e1e338u, 0x09d0889u, 0x5a9d7bdu, 0x00869f5u, 0xcb9f841u, 0xfcb79bcu, 0x14fcb3bu, 0xd78b460u, 0x9add89fu, 0x66e4a72u, 0xcdeb434u, 0x3157be7u, 0x5ed67e0u, 0x000140eu};
        return v[i];
    }
    FP_HD static constexpr uint32_t te_cy(int i) {   // f / s
        constexpr uint32_t v[14] = {0xcfd10d5u, 0x9457a38u, 0x383e475u, 0x7e96cb8u, 0x388e8d9u, 0xd00b5abu, 0xd32ebe1u, 0x0c15e03u, 0x5ac3c51u, 0xa270ad8u, 0xe68d2ebu, 0xbe54908u, 0x8029609u, 0x00001c4u};
        return v[i];
    }
    // Table entries of the integer accumulate path hold x 2^406 (the Montgomery form of fp29.cuh); plain numbers for the
    // conversions on the cold side:  mul(x 2^392, 2^406) = x 2^406,  mul(x 2^406, 2^378) = x 2^392.
    FP_HD static constexpr uint32_t te_c406(int i) {
        constexpr uint32_t v[14] = {0x9eaf730u, 0xe3ffffdu, 0x5c533e2u, 0x2718ec3u, 0xb6e2622u, 0xf5a2c7eu, 0x51ba717u, 0xb6edf83u, 0x9cbdd85u, 0xf055f3bu, 0x47628eau, 0xe55b0efu, 0x2a14aabu, 0x0001574u};
        return v[i];
    }
    FP_HD static constexpr uint32_t te_c378(int i) {
        constexpr uint32_t v[14] = {0xffffffeu, 0xee7ffffu, 0xffffef5u, 0x45779ffu, 0xfffd1e9u, 0xa08bed6u, 0x1c2193bu, 0xfe15d8eu, 0xbba4c19u, 0xbd6d89cu, 0x9f47f26u, 0xde2a738u, 0xb73d075u, 0x0000a38u};
        return v[i];
    }
};

// BLS12-377 with short-Weierstrass XYZZ buckets instead of the twisted-Edwards ones: the same field and curve, the group
// law complete on EVERY on-curve point (arkworks' contract: short_weierstrass_jacobian.rs:662-729 handles all of them),
// at 8 M + 2 S per mixed add instead of 7 M.  Selected per context (b200msm_set_bucket_arithmetic / B200MSM_GROUP=xyzz).
struct Bls12_377_sw : Bls12_377 {
    static constexpr bool HAS_TE = false;
    // arkworks multiplies the INTEGER scalar into the point (no reduction mod r): for a base outside G1, r P != 0, so neither
    // `s mod r` nor the top-bit fold is allowed.  This arithmetic walks all 256 scalar bits (+ 1 for the last signed carry).
    static constexpr int WINDOW_BITS = 257;
    static constexpr bool FOLD_SCALAR = false;
};

struct Bls12_381 {
    static constexpr int ID = 1;
    static constexpr int SCALAR_BITS = 255;
    static constexpr int WINDOW_BITS = 255;
    static constexpr bool FOLD_SCALAR = true;
    static constexpr uint32_t INV = 0xffcfffdu;
    static constexpr uint32_t P0INV = 0x0030003u;
    FP_HD static constexpr uint32_t p(int i) {
        constexpr uint32_t v[14] = {0xfffaaabu, 0xfefffffu, 0x3ffffb9u, 0xfffeb15u, 0x6241eabu, 0xa0f6b0fu, 0xf6730d2u,
                                    0xf38512bu, 0x4774b84u, 0x4bacd76u, 0xba7b643u, 0xe69a4b1u, 0x1ea397fu, 0x001a011u};
        return v[i];
    }
    FP_HD static constexpr uint32_t one(int i) {
        constexpr uint32_t v[14] = {0x347fcb8u, 0xd800000u, 0x002b119u, 0x0cde6d2u, 0xc7212e0u, 0x83a2090u, 0x037669fu,
                                    0xda0f73eu, 0x9b09b42u, 0x1297bb0u, 0x515d98fu, 0x012ca7cu, 0x659fcfau, 0x000577au};
        return v[i];
    }
    FP_HD static constexpr uint32_t cin(int i) {
        constexpr uint32_t v[14] = {0x80e6299u, 0x3500034u, 0xeb12856u, 0xdeb2699u, 0xc988670u, 0x4ef6697u, 0x70983e8u,
                                    0xa4e6fe9u, 0x3e8a053u, 0xecf271eu, 0xc20d323u, 0x6eb6385u, 0x47f1286u, 0x00156dau};
        return v[i];
    }
    FP_HD static constexpr uint32_t cout(int i) {
        constexpr uint32_t v[14] = {0x002fffdu, 0x0900000u, 0xc000276u, 0x000bc40u, 0x8baebf4u, 0x5753c75u, 0x55f4898u,
                                    0x7052574u, 0x7ce5853u, 0x56ec6d7u, 0x71a97a2u, 0xe4935c0u, 0xec3fa80u, 0x0015f65u};
        return v[i];
    }
    static constexpr bool HAS_TE = false;            // cofactor not divisible by 4: no Montgomery / twisted-Edwards model
    FP_HD static constexpr uint32_t te_k(int) { return 0; }
    FP_HD static constexpr uint32_t te_f(int) { return 0; }
    FP_HD static constexpr uint32_t te_s(int) { return 0; }
    FP_HD static constexpr uint32_t te_s3(int) { return 0; }
    FP_HD static constexpr uint32_t te_cy(int) { return 0; }
    FP_HD static constexpr uint32_t te_c406(int) { return 0; }
    FP_HD static constexpr uint32_t te_c378(int) { return 0; }
    FP_HD static constexpr uint32_t r(int i) {
        constexpr uint32_t v[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return v[i];
    }
};

// limb i of K*p in limb-dominant form: limbs 0..12 of K*p each borrow 2^28 from the limb above, so every limb is
// >= 2^28 - 1 while the value is still exactly K*p.
template <class C> FP_HD constexpr uint32_t kp_dom(int k, int i) {
    uint64_t carry = 0;
    uint32_t limb = 0;
    for (int j = 0; j <= i; j++) {
        uint64_t t = (uint64_t)C::p(j) * (uint32_t)k + carry;
        limb = (uint32_t)(t & FP_MASK);
        carry = t >> FP_BITS;
        if (j == FP_LIMBS - 1) limb = (uint32_t)t;      // top limb keeps everything
    }
    if (i < FP_LIMBS - 1) limb += (1u << FP_BITS);      // borrowed from above
    if (i > 0) limb -= 1;                               // lent to below
    return limb;
}

// Keeps a value opaque to the optimiser as a 32-bit register.  Without it NVVM widens the Montgomery factor m to
// 64 bits (neg.s64 / and.b64) and every m * p_j becomes a 64 x 64 mul.lo.s64 -- one junk IADD3 per product in SASS.
FP_HD uint32_t opaque32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    asm("" : "+r"(v));
#endif
    return v;
}

template <class C> struct Fp {
    static constexpr int N = FP_LIMBS;

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
    // carry-propagate so limbs 0..12 are below 2^28 (limbs may be up to 2^32 - 1 on entry)
    FP_HD static void norm(fp &a) {
        uint32_t c = 0;
#pragma unroll
        for (int i = 0; i < N - 1; i++) {
            uint32_t t = a.l[i] + c;
            // a.l[i] + c cannot wrap: callers keep limbs below 2^31 and c < 2^4
            a.l[i] = t & FP_MASK;
            c = t >> FP_BITS;
        }
        a.l[N - 1] += c;
    }
    // r = a + b, normalised
    FP_HD static void add(fp &r, const fp &a, const fp &b) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = a.l[i] + b.l[i];
        norm(r);
    }
    // r = a + K*p - b, normalised.  Requires b normalised, value(b) < (K-1) p, limbs of a below 2^30.
    template <int K> FP_HD static void sub(fp &r, const fp &a, const fp &b) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = a.l[i] + kp_dom<C>(K, i) - b.l[i];
        norm(r);
    }
    // The same two without the carry propagation: limbs grow (add: sum of the operands' limb bounds; sub: a's + 2^29), the
    // value is unchanged.  The multipliers accept such operands within the column bounds derived at Te::madd (te.cuh).
    FP_HD static void add_raw(fp &r, const fp &a, const fp &b) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = a.l[i] + b.l[i];
    }
    template <int K> FP_HD static void sub_raw(fp &r, const fp &a, const fp &b) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = a.l[i] + kp_dom<C>(K, i) - b.l[i];
    }
    // r = K*p - a (a normalised, value(a) < (K-1) p), normalised
    template <int K> FP_HD static void neg(fp &r, const fp &a) {
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = kp_dom<C>(K, i) - a.l[i];
        norm(r);
    }

    // Montgomery product a*b / 2^392 mod p (normalised, value < 2p under the bounds in the header comment).
    // Row i: t += a * b_i (14 IMAD.WIDE, no carries);  m = -t_0 / p mod 2^28;  t += m * p  (14 more; p_0 = 1 on BLS12-377
    // costs nothing);  t_0 is now a multiple of 2^28: push t_0 >> 28 into t_1 and rename the columns down by one.
    FP_HD static void mul_schoolbook(fp &r, const fp &a, const fp &b) {
        uint64_t t[N];
#pragma unroll
        for (int i = 0; i < N; i++) {
            const uint32_t bi = b.l[i];
            // the top column is fresh every round (written, not accumulated: no zero-initialised registers to add)
#pragma unroll
            for (int j = 0; j < N; j++) {
                if (i == 0 || j == N - 1) t[j] = (uint64_t)a.l[j] * bi;
                else t[j] += (uint64_t)a.l[j] * bi;
            }
            const uint32_t m = opaque32(((uint32_t)t[0] * C::INV) & FP_MASK);
#pragma unroll
            for (int j = 0; j < N; j++) t[j] += (uint64_t)m * C::p(j);
            const uint64_t c = t[0] >> FP_BITS;
#pragma unroll
            for (int j = 0; j < N - 1; j++) t[j] = t[j + 1];
            t[0] += c;
        }
        // columns 0..12 hold the result (column 13 of the last round was shifted into t[12])
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < N - 2; j++) {
            uint64_t v = t[j] + c;
            r.l[j] = (uint32_t)v & FP_MASK;
            c = v >> FP_BITS;
        }
        {
            uint64_t v = t[N - 2] + c;
            r.l[N - 2] = (uint32_t)v & FP_MASK;
            r.l[N - 1] = (uint32_t)(v >> FP_BITS);
        }
    }
    // The same product with one level of Karatsuba on a * b (a = a0 + a1 2^196, 7 limbs each):
    //   z0 = a0 b0, z2 = a1 b1, z1 = (a0 + a1)(b0 + b1) - z0 - z2   -- 3 x 49 = 147 IMAD.WIDE instead of 196,
    // then the 14 Montgomery rounds on the 28 columns (182): 329 against 378.  (A second level, 7 = 4 + 3, reaches 305
    // but its extra live columns cost more in spills than the 24 multiplies save: 235 vs 228 ms at 2^26.)  The kernel is bound by the integer
    // multiply pipe (32 IMAD.WIDE lanes/clk/SM) while the ALU pipe is a quarter busy, so trading 49 multiplies for ~90
    // 64-bit adds is a net gain.  Needs NORMALISED operands (limbs 0..12 below 2^28, top limb below 2^27): the half
    // sums are then below 2^29, a middle column below 7 * 2^58, and every identity holds per column (all column values
    // are sums of non-negative products, so z1's columns never go negative).
    FP_HD static void mul(fp &r, const fp &a, const fp &b) { mul_karatsuba(r, a, b); }
    FP_HD static void mul_karatsuba(fp &r, const fp &a, const fp &b) {
        constexpr int H = N / 2;      // 7
        uint64_t t[2 * N];
        uint64_t mid[2 * H - 1];
        uint32_t sa[H], sb[H];
#pragma unroll
        for (int i = 0; i < H; i++) { sa[i] = a.l[i] + a.l[H + i]; sb[i] = b.l[i] + b.l[H + i]; }
#pragma unroll
        for (int k = 0; k < 2 * N; k++) t[k] = 0;
        // z0 -> columns 0..12, z2 -> columns 14..26, middle product -> mid
#pragma unroll
        for (int i = 0; i < H; i++) {
#pragma unroll
            for (int j = 0; j < H; j++) {
                if (i == 0 || j == H - 1) {
                    t[i + j] = (uint64_t)a.l[i] * b.l[j];
                    t[N + i + j] = (uint64_t)a.l[H + i] * b.l[H + j];
                    mid[i + j] = (uint64_t)sa[i] * sb[j];
                } else {
                    t[i + j] += (uint64_t)a.l[i] * b.l[j];
                    t[N + i + j] += (uint64_t)a.l[H + i] * b.l[H + j];
                    mid[i + j] += (uint64_t)sa[i] * sb[j];
                }
            }
        }
        // z1 = mid - z0 - z2 (per column), added at column offset 7
#pragma unroll
        for (int k = 0; k < 2 * H - 1; k++) mid[k] -= t[k] + t[N + k];
#pragma unroll
        for (int k = 0; k < 2 * H - 1; k++) t[H + k] += mid[k];
        // Montgomery reduction of the 28 columns
#pragma unroll
        for (int i = 0; i < N; i++) {
            const uint32_t m = opaque32(((uint32_t)t[i] * C::INV) & FP_MASK);
#pragma unroll
            for (int j = 0; j < N; j++) t[i + j] += (uint64_t)m * C::p(j);
            t[i + 1] += t[i] >> FP_BITS;
        }
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < N - 1; j++) {
            uint64_t v = t[N + j] + c;
            r.l[j] = (uint32_t)v & FP_MASK;
            c = v >> FP_BITS;
        }
        r.l[N - 1] = (uint32_t)(t[2 * N - 1] + c);
    }
    // a^2: the 91 cross products are taken once against the doubled operand (limbs of a below 2^29 here)
    FP_HD static void sqr(fp &r, const fp &a) {
        uint64_t t[2 * N];
#pragma unroll
        for (int j = 0; j < 2 * N; j++) t[j] = 0;
        uint32_t a2[N];
#pragma unroll
        for (int j = 0; j < N; j++) a2[j] = a.l[j] << 1;
#pragma unroll
        for (int i = 0; i < N; i++) {
            t[2 * i] += (uint64_t)a.l[i] * a.l[i];
#pragma unroll
            for (int j = i + 1; j < N; j++) t[i + j] += (uint64_t)a.l[i] * a2[j];
        }
        // Montgomery reduction of the 28 columns: column i is cleared, then its carry moves up
#pragma unroll
        for (int i = 0; i < N; i++) {
            const uint32_t m = opaque32(((uint32_t)t[i] * C::INV) & FP_MASK);
#pragma unroll
            for (int j = 0; j < N; j++) t[i + j] += (uint64_t)m * C::p(j);
            t[i + 1] += t[i] >> FP_BITS;
        }
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < N - 1; j++) {
            uint64_t v = t[N + j] + c;
            r.l[j] = (uint32_t)v & FP_MASK;
            c = v >> FP_BITS;
        }
        r.l[N - 1] = (uint32_t)(t[2 * N - 1] + c);
    }

    // canonical representative in [0, p): one multiply by the Montgomery one (value < 2p), then a conditional subtract
    FP_HD static void reduce(fp &a) {
        fp one;
        set_one(one);
        fp t;
        mul(t, a, one);
        csub_p(t);
        a = t;
    }
    // a normalised, value < 2p:  if (a >= p) a -= p
    FP_HD static void csub_p(fp &a) {
        uint32_t d[N];
        uint32_t borrow = 0;
#pragma unroll
        for (int i = 0; i < N; i++) {
            uint32_t t = a.l[i] - C::p(i) - borrow;
            borrow = t >> 31;
            d[i] = (i < N - 1) ? (t & FP_MASK) : t;
        }
#pragma unroll
        for (int i = 0; i < N; i++) a.l[i] = borrow ? a.l[i] : d[i];
    }
    // a == 0 (mod p) for a lazily reduced a (value below 2^15 p).  A multiple k*p has low limb k*p_0, so
    // k = a_0 / p_0 mod 2^28 must be small: one multiply filters all but 2^-12 of the non-zero values, the rest take
    // the exact path.
    FP_HD static bool is_zero(const fp &a) {
        uint32_t k = (a.l[0] * C::P0INV) & FP_MASK;
        if (k >= (1u << 16)) return false;
        fp t = a;
        norm(t);
        reduce(t);
        return is_literal_zero(t);
    }

    // arkworks ABI (6 x u64, R = 2^384, canonical) <-> internal (14 x 28, R = 2^392, canonical)
    FP_HD static void repack_from_abi(fp &r, const fp_abi &a) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            const int bit = FP_BITS * i, w = bit >> 5, s = bit & 31;
            uint64_t v = a.w[w];
            if (w + 1 < 12) v |= (uint64_t)a.w[w + 1] << 32;
            r.l[i] = (uint32_t)(v >> s) & FP_MASK;
        }
    }
    FP_HD static void repack_to_abi(fp_abi &r, const fp &a) {
#pragma unroll
        for (int w = 0; w < 12; w++) {
            // bits [32w, 32w+32) come from at most three 28-bit limbs
            const int bit = 32 * w, i = bit / FP_BITS, s = bit % FP_BITS;
            uint64_t v = (uint64_t)a.l[i] >> s;
            if (i + 1 < N) v |= (uint64_t)a.l[i + 1] << (FP_BITS - s);
            if (i + 2 < N && 2 * FP_BITS - s < 32) v |= (uint64_t)a.l[i + 2] << (2 * FP_BITS - s);
            r.w[w] = (uint32_t)v;
        }
    }
    FP_HD static void from_abi(fp &r, const fp_abi &a) {
        fp t, c;
        repack_from_abi(t, a);
#pragma unroll
        for (int i = 0; i < N; i++) c.l[i] = C::cin(i);
        mul(r, t, c);
        csub_p(r);
    }
    FP_HD static void to_abi(fp_abi &r, const fp &a) {
        fp t = a, c;
        norm(t);
#pragma unroll
        for (int i = 0; i < N; i++) c.l[i] = C::cout(i);
        mul(t, t, c);
        csub_p(t);
        repack_to_abi(r, t);
    }

    // limb i of p - 2 (exponent of the Fermat inversion), 28-bit limbs
    FP_HD static constexpr uint32_t pm2_const(int i) {
        uint32_t borrow = 2, out = 0;
        for (int j = 0; j <= i; j++) {
            uint32_t pj = C::p(j);
            out = (pj - borrow) & FP_MASK;
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
    // a^(p-2) = 1/a (a normalised, value < 2p; result normalised, < 2p)
    FP_HD static void inv(fp &r, const fp &a) {
        fp acc;
        set_one(acc);
        for (int i = N * FP_BITS - 1; i >= 0; i--) {
            mul(acc, acc, acc);
            if ((pm2_limb(i / FP_BITS) >> (i % FP_BITS)) & 1) mul(acc, acc, a);
        }
        r = acc;
    }
};

}  // namespace b200msm

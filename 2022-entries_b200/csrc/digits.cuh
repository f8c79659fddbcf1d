// digits.cuh -- a4: scalar -> signed window digits -> (bucket key, point reference) pairs; the arithmetic of k_digits and
// k_digits_bin (kernels_hot.cuh has the description of the scheme).  Plain C++ apart from the markers below, so
// tests/host_emul builds it with g++ and checks the decomposition against Python integers without a GPU.
#pragma once
#include "fp.cuh"

#if defined(__CUDACC__)
#define DG_D __device__ __forceinline__
#define DG_LDG(p) __ldg(p)
#else
#define DG_D inline
#define DG_LDG(p) (*(p))
#ifndef __VECTOR_TYPES_H__
struct uint4 { uint32_t x, y, z, w; };
#endif
#endif

namespace b200msm {

template <class C> struct ScalarOps {
    // s >= k*r ?  s -= k*r   (k*r must fit 256 bits)
    template <int K> DG_D static void csub_kr(uint32_t (&s)[8]) {
        uint32_t kr[8];
        uint64_t carry = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { uint64_t t = (uint64_t)C::r(i) * K + carry; kr[i] = (uint32_t)t; carry = t >> 32; }
        uint32_t t[8];
        uint32_t borrow = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t d = (uint64_t)s[i] - kr[i] - borrow;
            t[i] = (uint32_t)d;
            borrow = (uint32_t)(d >> 63);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = borrow ? s[i] : t[i];
    }
    DG_D static void mod_r(uint32_t (&s)[8]) {
        if (C::SCALAR_BITS <= 253) { csub_kr<8>(s); csub_kr<4>(s); }
        csub_kr<2>(s);
        csub_kr<1>(s);
    }
    // returns true when s was replaced by r - s
    DG_D static bool fold_top_bit(uint32_t (&s)[8]) {
        const int tb = C::SCALAR_BITS - 1;
        if (!((s[tb >> 5] >> (tb & 31)) & 1)) return false;
        uint32_t borrow = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t d = (uint64_t)C::r(i) - s[i] - borrow;
            s[i] = (uint32_t)d;
            borrow = (uint32_t)(d >> 63);
        }
        return true;
    }
};

// walks the signed digits of scalar i (the scalar of base number gi): fn(bucket key, point reference) for every
// non-zero digit
template <class C, int MAXW = 0, class FN>
DG_D void for_each_digit(const uint32_t *__restrict__ scalars, uint32_t i, uint32_t gi, uint32_t table_stride, int c, int W, int F, FN fn) {
    uint32_t s[9];
    {
        const uint4 *p = reinterpret_cast<const uint4 *>(scalars) + 2 * (size_t)i;
        uint4 lo = DG_LDG(p), hi = DG_LDG(p + 1);
        s[0] = lo.x; s[1] = lo.y; s[2] = lo.z; s[3] = lo.w; s[4] = hi.x; s[5] = hi.y; s[6] = hi.z; s[7] = hi.w;
    }
    uint32_t(&s8)[8] = reinterpret_cast<uint32_t(&)[8]>(s);
    bool neg = false;
    if (C::FOLD_SCALAR) {
        if (s[7] >= C::r(7)) ScalarOps<C>::mod_r(s8);  // top word below r's: already reduced (every harness scalar is)
        neg = ScalarOps<C>::fold_top_bit(s8);
    }
    s[8] = 0;
    const uint32_t half = 1u << (c - 1), mask = (1u << c) - 1;
    uint32_t carry = 0;
#pragma unroll
    for (int w = 0; w < (MAXW > 0 ? MAXW : W); w++) {       // MAXW > 0: compile-time trip count, so fn may index registers by w
        if (MAXW > 0 && w >= W) break;
        int pos = w * c;
        int wi = pos >> 5, sh = pos & 31;
        uint64_t v = (uint64_t)s[wi] | ((uint64_t)s[wi + 1 > 8 ? 8 : wi + 1] << 32);
        uint32_t d = ((uint32_t)(v >> sh) & mask) + carry;
        bool flip = d > half;
        if (flip) d = (1u << c) - d;
        carry = flip ? 1u : 0u;
        if (d == 0) continue;
        if constexpr (MAXW > 0) fn(w, (uint32_t)(w % F) * half + d - 1, ((uint32_t)(w / F) * table_stride + gi) | ((neg != flip) ? 0x80000000u : 0u));
        else fn((uint32_t)(w % F) * half + d - 1, ((uint32_t)(w / F) * table_stride + gi) | ((neg != flip) ? 0x80000000u : 0u));
    }
}

}  // namespace b200msm

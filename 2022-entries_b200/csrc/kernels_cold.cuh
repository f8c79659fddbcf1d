// kernels_cold.cuh -- kernels that run once per MSM on small data or once per context (templated on the curve):
//   k_reduce_slices  a9  (sum, weighted sum) of m consecutive buckets per thread; applied level by level
//   k_finalize       a10 last <=32-element sums, window weights, Horner over folded windows, normalisation
//   k_double_n / k_batch_normalize / k_gen_*   a11  base tables 2^(c*F*j) * P_i, synthetic bases and scalars
//   k_test_*         element-wise field / curve operations for the parity tests
// Curve operations here go through the non-inlined Ec<C>::*_ni entry points: these kernels are latency- or
// launch-bound, and inlining ~14 Montgomery multiplies at every call site only costs compile time and I-cache.
#pragma once
#include "kernels_util.cuh"
#include "launch.h"

namespace b200msm {

// ---------------------------------------------------------------------------------------------
// a9: bucket reduction.  For a window with buckets B_0..B_{n-1} (bucket k carries weight k+1) the window sum is
//   sum_k B_k + W0(B),  W0(B) = sum_k k*B_k = sum_t T_t + m * W0(S),   S_t = sum_j B_{tm+j},  T_t = sum_j j*B_{tm+j}.
// One thread produces (S_t, T_t) for m consecutive inputs with the running-sum trick (2 adds per input); the host
// applies the kernel level by level until 32 or fewer elements are left (k_finalize).  INDIRECT reads level 0
// through the piece table (a bucket is the sum of its pieces, almost always exactly one).  SUM_ONLY skips T.
template <class C, bool INDIRECT, bool SUM_ONLY>
__global__ void __launch_bounds__(128) k_reduce_slices(const xyzz_t *__restrict__ in, const uint32_t *__restrict__ piece_off, uint32_t n_out, uint32_t m,
                                                       xyzz_t *__restrict__ outS, xyzz_t *__restrict__ outT) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_out) return;
    xyzz_t run, tot, b;
    Ec<C>::set_inf(run);
    Ec<C>::set_inf(tot);
    for (uint32_t jj = m; jj-- > 0;) {
        uint32_t k = t * m + jj;
        if (!SUM_ONLY && jj + 1 < m) Ec<C>::add_ni(tot, run);      // after the loop: tot = sum_j j * B_j
        if (INDIRECT) {
            uint32_t p0 = piece_off[k], p1 = piece_off[k + 1];
            for (uint32_t q = p0; q < p1; q++) { load_xyzz(b, in + q); Ec<C>::add_ni(run, b); }
        } else {
            load_xyzz(b, in + k);
            Ec<C>::add_ni(run, b);
        }
    }
    store_xyzz(outS + t, run);
    if (!SUM_ONLY) store_xyzz(outT + t, tot);
}

// ---------------------------------------------------------------------------------------------
// a10: finalize (one warp).  Per folded window f the host hands a list of terms, each an array of <= 32 points:
//   kind 0: 2^shift * sum(arr)        kind 1: 2^shift * sum_k k * arr[k]
// window value = sum of its terms; result = Horner over windows with c doublings; then either the XYZZ value is
// stored (partial result of one GPU) or it is normalised to the caller's Jacobian (x, y, z = 1) / (1, 1, 0).

template <class C> __device__ void warp_sum(xyzz_t &v, xyzz_t *sh) {
    // tree sum of the 32 lanes' points through shared memory; result in lane 0
    const int lane = threadIdx.x & 31;
    for (int o = 16; o > 0; o >>= 1) {
        __syncwarp();
        if (lane >= o && lane < 2 * o) store_xyzz(sh + lane - o, v);
        __syncwarp();
        if (lane < o) { xyzz_t t; load_xyzz(t, sh + lane); Ec<C>::add_ni(v, t); }
    }
    __syncwarp();
}
template <class C>
__global__ void __launch_bounds__(32) k_finalize(fin_args_t a, xyzz_t *__restrict__ out_xyzz, uint32_t *__restrict__ out_jac, int extra_n, const xyzz_t *__restrict__ extra) {
    __shared__ xyzz_t sh[16];
    const int lane = threadIdx.x;
    xyzz_t total;
    Ec<C>::set_inf(total);
    for (int f = a.F - 1; f >= 0; f--) {
        xyzz_t win;
        Ec<C>::set_inf(win);
        for (int ti = 0; ti < a.nterms; ti++) {
            const fin_term_t &tm = a.term[ti];
            xyzz_t v;
            Ec<C>::set_inf(v);
            if ((uint32_t)lane < tm.n) {
                load_xyzz(v, tm.arr + (size_t)f * tm.stride_f + lane);
                if (tm.kind == 1) {         // lane * v by double-and-add over 5 bits
                    xyzz_t r;
                    Ec<C>::set_inf(r);
                    for (int b = 4; b >= 0; b--) { Ec<C>::dbl_ni(r); if ((lane >> b) & 1) Ec<C>::add_ni(r, v); }
                    v = r;
                }
            }
            warp_sum<C>(v, sh);
            if (lane == 0) {
                for (uint32_t s = 0; s < tm.shift; s++) Ec<C>::dbl_ni(v);
                Ec<C>::add_ni(win, v);
            }
        }
        if (lane == 0) {
            if (f != a.F - 1) for (int s = 0; s < a.c; s++) Ec<C>::dbl_ni(total);
            Ec<C>::add_ni(total, win);
        }
    }
    // optional: add further partial results (other GPUs' XYZZ partials, lane-parallel)
    if (extra_n > 0) {
        xyzz_t v;
        Ec<C>::set_inf(v);
        for (int k = lane; k < extra_n; k += 32) { xyzz_t t; load_xyzz(t, extra + k); Ec<C>::add_ni(v, t); }
        warp_sum<C>(v, sh);
        if (lane == 0) Ec<C>::add_ni(total, v);
    }
    if (lane == 0) {
        if (out_xyzz) {
            Fp<C>::reduce(total.x); Fp<C>::reduce(total.y); Fp<C>::reduce(total.zz); Fp<C>::reduce(total.zzz);
            store_xyzz(out_xyzz, total);
        }
        if (out_jac) {
            // arkworks G1Projective (Jacobian x, y, z; 3 x 48 bytes): finite -> (x_aff, y_aff, 1); infinity -> (1, 1, 0)
            fp x, y, z;
            if (Ec<C>::is_inf(total)) { Fp<C>::set_one(x); Fp<C>::set_one(y); Fp<C>::set_zero(z); }
            else { Ec<C>::to_affine_ni(x, y, total); Fp<C>::set_one(z); }
#pragma unroll
            for (int i = 0; i < 12; i++) { out_jac[i] = x.l[i]; out_jac[12 + i] = y.l[i]; out_jac[24 + i] = z.l[i]; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// a11: base tables.
// tmp[i] = 2^k * src[i]  (XYZZ, not normalised)
template <class C>
__global__ void __launch_bounds__(128) k_double_n(const affine_t *__restrict__ src, uint32_t n, int k, xyzz_t *__restrict__ tmp) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine_t a;
    load_affine(a, src + i);
    xyzz_t q;
    Ec<C>::dbl_affine_ni(q, a.x, a.y);
    for (int s = 1; s < k; s++) Ec<C>::dbl_ni(q);
    store_xyzz(tmp + i, q);
}
// XYZZ -> canonical affine with one inversion per BATCH points (Montgomery's trick on ZZZ).
// A ZZZ of zero (only reachable from bases outside the prime-order subgroup) is replaced by 1 so it cannot poison
// its batch; that entry's output is meaningless, as the contract (bases in G1) allows.
template <class C>
__global__ void __launch_bounds__(128) k_batch_normalize(const xyzz_t *__restrict__ tmp, uint32_t n, affine_t *__restrict__ dst) {
    typedef Fp<C> F;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t base = t * NORM_BATCH;
    if (base >= n) return;
    uint32_t cnt = n - base < NORM_BATCH ? n - base : NORM_BATCH;
    fp prefix[NORM_BATCH];
    fp acc;
    F::set_one(acc);
    for (uint32_t k = 0; k < cnt; k++) {
        fp z;
        const uint4 *s = reinterpret_cast<const uint4 *>(&tmp[base + k].zzz);
        uint4 *d = reinterpret_cast<uint4 *>(&z);
        d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
        if (F::is_zero(z)) F::set_one(z);
        prefix[k] = acc;
        Ec<C>::mul_ni(acc, acc, z);
    }
    fp inv;
    Ec<C>::inv_ni(inv, acc);
    for (uint32_t k = cnt; k-- > 0;) {
        xyzz_t q;
        load_xyzz(q, tmp + base + k);
        if (F::is_zero(q.zzz)) F::set_one(q.zzz);
        fp i1;
        Ec<C>::mul_ni(i1, inv, prefix[k]);      // 1 / ZZZ_k
        Ec<C>::mul_ni(inv, inv, q.zzz);         // drop ZZZ_k from the running inverse
        affine_t o;
        fp i2, z2;
        Ec<C>::mul_ni(i2, i1, i1);
        Ec<C>::mul_ni(z2, q.zz, q.zz);
        Ec<C>::mul_ni(o.x, q.x, z2);
        Ec<C>::mul_ni(o.x, o.x, i2);
        Ec<C>::mul_ni(o.y, q.y, i1);
        F::reduce(o.x);
        F::reduce(o.y);
        store_affine(dst + base + k, o);
    }
}
// test / benchmark bases with known discrete logs: tmp[i] = k_i * G
//   mode 0 ("harness"): k_i = splitmix64(seed + (i mod distinct)) | 1   -- `distinct` random points replicated, the
//                        shape of the reference generator (util.rs:15-28: 2^15 points, vector doubled up to N)
//   mode 1 ("distinct"): k_i = seed + i                                   -- N different points
template <class C>
__global__ void __launch_bounds__(128) k_gen_points(uint32_t first, uint32_t n, uint64_t seed, int mode, uint32_t distinct, fp gx, fp gy, xyzz_t *__restrict__ tmp) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint32_t i = first + t;
    uint64_t k = mode == 0 ? (splitmix64(seed + (i % distinct)) | 1ull) : seed + i;
    xyzz_t q;
    Ec<C>::set_inf(q);
    for (int b = 63; b >= 0; b--) {
        Ec<C>::dbl_ni(q);
        if ((k >> b) & 1) Ec<C>::madd_ni(q, gx, gy);
    }
    store_xyzz(tmp + t, q);
}
// uniform scalars in [0, r): counter-based SplitMix64 stream, top limb masked to SCALAR_BITS, rejection
template <class C>
__global__ void __launch_bounds__(256) k_gen_scalars(uint32_t n, uint64_t seed, uint32_t *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    for (uint32_t attempt = 0;; attempt++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint64_t v = splitmix64(seed ^ (((uint64_t)i * 4 + k) * 0x100 + attempt));
            s[2 * k] = (uint32_t)v;
            s[2 * k + 1] = (uint32_t)(v >> 32);
        }
        s[7] &= (C::SCALAR_BITS & 31) ? ((1u << (C::SCALAR_BITS & 31)) - 1) : 0xffffffffu;
        bool lt = false;   // s < r ?
        for (int k = 7; k >= 0; k--) {
            if (s[k] != C::r(k)) { lt = s[k] < C::r(k); break; }
        }
        if (lt) break;
    }
    uint4 *o = reinterpret_cast<uint4 *>(out) + 2 * (size_t)i;
    o[0] = make_uint4(s[0], s[1], s[2], s[3]);
    o[1] = make_uint4(s[4], s[5], s[6], s[7]);
}
// test entry points: element-wise field / curve operations on device arrays
template <class C>
__global__ void k_test_fp(int op, const fp *a, const fp *b, fp *r, uint32_t n) {
    typedef Fp<C> F;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp x = a[i], y = b[i], t;
    switch (op) {
    case 0: F::mul(t, x, y); break;
    case 1: F::add(t, x, y); break;
    case 2: F::template sub<2>(t, x, y); break;
    case 3: F::sqr(t, x); break;
    case 4: Ec<C>::inv_ni(t, x); break;
    case 5: F::to_mont(t, x); break;
    case 6: F::from_mont(t, x); break;
    default: t = x;
    }
    F::reduce(t);
    r[i] = t;
}
template <class C>
__global__ void k_test_ec(int op, xyzz_t *acc, const void *other, uint32_t n, int reps) {
    typedef Ec<C> E;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xyzz_t q;
    load_xyzz(q, acc + i);
    for (int k = 0; k < reps; k++) {
        if (op == 0) { affine_t a; load_affine(a, (const affine_t *)other + i); E::madd(q, a.x, a.y); }
        else if (op == 1) { xyzz_t o; load_xyzz(o, (const xyzz_t *)other + i); E::add_ni(q, o); }
        else if (op == 2) E::dbl_ni(q);
        else if (op == 3) { affine_t a; load_affine(a, (const affine_t *)other + i); E::dbl_affine_ni(q, a.x, a.y); }
    }
    Fp<C>::reduce(q.x); Fp<C>::reduce(q.y); Fp<C>::reduce(q.zz); Fp<C>::reduce(q.zzz);
    store_xyzz(acc + i, q);
}

}  // namespace b200msm

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
// One thread produces (S_t, T_t) for m consecutive inputs with the running-sum trick (2 adds per input); S_t is
// stored, the T_t of a block are tree-summed in shared memory (only their plain sum is ever needed) and stored once
// per block.  The host applies the kernel level by level (m = 8: every level but the first is latency-bound, so
// short serial chains and many threads) until 32 or fewer S are left for k_finalize.  INDIRECT reads level 0 through
// the piece table (a bucket is the sum of its pieces, almost always exactly one).
template <class C> __device__ __forceinline__ void block_tree_sum(xyzz_t &v, xyzz_t *sh) {
    for (uint32_t o = blockDim.x >> 1; o > 0; o >>= 1) {
        __syncthreads();
        if (threadIdx.x >= o && threadIdx.x < 2 * o) store_xyzz(sh + threadIdx.x - o, v);
        __syncthreads();
        if (threadIdx.x < o) { xyzz_t t; load_xyzz(t, sh + threadIdx.x); Grp<C>::add_ni(v, t); }
    }
}
template <class C> __device__ void warp_sum(xyzz_t &v, xyzz_t *sh);
// A bucket split into many pieces is summed by the whole warp -- lanes take pieces 32 apart, then a shared-memory tree --
// instead of serially by the one lane that owns the bucket.  Two sources of such buckets: skewed scalar sets (all-equal,
// tiny or sparse scalars put millions of points into a handful of buckets), and a top window of only a few bits (c not
// dividing the scalar length: e.g. c = 21 leaves a 1-bit top window whose single bucket receives n/2 points).  Per warp
// the cheaper of the two walks is chosen: serial costs max(pieces) adds, cooperative costs sum(pieces/32 + 6) over the
// lanes with more than BIG_BUCKET_PIECES pieces.  Uniform scalars with a full top window never enter this path.
#define BIG_BUCKET_PIECES 8
// Multi-GPU bucket reduce-scatter (SURVEY.md 8e asks for one exchange of partial sums; this is the exchange done where the
// data is consumed): with point sharding every GPU holds partial sums of ALL buckets.  GPU g reduces only the buckets
// [g nb / G, (g + 1) nb / G), reading every peer's partial sums of that slice straight from the peer's HBM over NVLink
// (mapped pointers in srcs) inside this kernel -- the loads of one thread's next source overlap the curve additions of the
// other warps, there is no staging copy and no separate collective.  Cross-GPU ordering is two counters per pair:
// `ready` (my bucket sums of epoch e are complete; set by k_peer_signal after k_accumulate) is awaited here before the first
// remote load, `done` (I have finished reading yours) is awaited by the peer before it overwrites them (k_peer_wait).
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// peer = true: 16-byte loads that fetch again (ld.global.cv): the record lives in a peer GPU's memory and was written while
// this launch was already waiting; peer = false: this GPU's own data, complete before the launch -- ordinary cached loads
__device__ __forceinline__ void load_xyzz_src(xyzz_t &d, const xyzz_t *src, bool peer) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *t = reinterpret_cast<uint4 *>(&d);
    if (peer) {
#pragma unroll
        for (int i = 0; i < XYZZ_VEC; i++) t[i] = __ldcv(s + i);
    } else {
#pragma unroll
        for (int i = 0; i < XYZZ_VEC; i++) t[i] = s[i];
    }
}
template <class C, bool INDIRECT>
__global__ void __launch_bounds__(128, 3) k_reduce_slices(const xyzz_t *__restrict__ in, const red_src_t *__restrict__ srcs, uint32_t nsrc, uint32_t bucket_base, uint32_t m,
                                                       xyzz_t *__restrict__ outS, xyzz_t *__restrict__ outT,
                                                       const uint32_t *__restrict__ wait_flags, uint32_t wait_n, uint32_t epoch) {
    __shared__ xyzz_t sh[64];
    __shared__ xyzz_t shw[INDIRECT ? 4 : 1][INDIRECT ? 17 : 1];     // per warp: 16 tree slots + 1 hand-over slot
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;      // the grid covers the outputs exactly
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (INDIRECT && wait_flags) {                                   // peers' bucket sums of this epoch must be complete
        if (threadIdx.x < wait_n)
            while ((int32_t)(ld_acquire_sys(wait_flags + threadIdx.x) - epoch) < 0) __nanosleep(100);
        __syncthreads();
    }
    xyzz_t run, tot, b;
    Grp<C>::set_inf(run);
    Grp<C>::set_inf(tot);
    for (uint32_t jj = m; jj-- > 0;) {
        uint32_t k = bucket_base + t * m + jj;
        if (jj + 1 < m) Grp<C>::add_ni(tot, run);                   // after the loop: tot = sum_j j * B_j
        if (INDIRECT) {        // a bucket = the sum of its pieces, over every chunk the point set was accumulated in, on every GPU
            for (uint32_t ch = 0; ch < nsrc; ch++) {
                const uint32_t *po = srcs[ch].poff;
                const bool peer = wait_flags != nullptr;
                const uint32_t p0 = peer ? __ldcv(po + k) : po[k], p1 = peer ? __ldcv(po + k + 1) : po[k + 1];
                const xyzz_t *src = srcs[ch].res;
                bool big = (blockDim.x & 31) == 0 && p1 - p0 > BIG_BUCKET_PIECES;            // whole warps only
                uint32_t todo = __ballot_sync(0xffffffffu, big);
                if (todo) {                                         // serial: max(pieces); cooperative: sum(pieces / 32 + 6)
                    uint32_t serial = p1 - p0, coop = big ? (p1 - p0) / 32 + 6 : 0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const uint32_t t2 = __shfl_xor_sync(0xffffffffu, serial, o);
                        serial = t2 > serial ? t2 : serial;
                        coop += __shfl_xor_sync(0xffffffffu, coop, o);
                    }
                    if (coop >= serial) { todo = 0; big = false; }
                }
                while (todo) {                                      // warp-uniform: every lane walks the same list
                    const int owner = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const uint32_t q0 = __shfl_sync(0xffffffffu, p0, owner), q1 = __shfl_sync(0xffffffffu, p1, owner);
                    xyzz_t part;
                    Grp<C>::set_inf(part);
                    for (uint32_t q = q0 + lane; q < q1; q += 32) { load_xyzz_src(b, src + q, peer); Grp<C>::add_ni(part, b); }
                    warp_sum<C>(part, shw[warp]);
                    if (lane == 0) store_xyzz(&shw[warp][16], part);
                    __syncwarp();
                    if ((int)lane == owner) { load_xyzz(b, &shw[warp][16]); Grp<C>::add_ni(run, b); }
                    __syncwarp();
                }
                if (!big) for (uint32_t q = p0; q < p1; q++) { load_xyzz_src(b, src + q, peer); Grp<C>::add_ni(run, b); }
            }
        } else {
            load_xyzz(b, in + k);
            Grp<C>::add_ni(run, b);
        }
    }
    store_xyzz(outS + t, run);
    block_tree_sum<C>(tot, sh);
    if (threadIdx.x == 0) store_xyzz(outT + blockIdx.x, tot);
}
// Multi-GPU bucket reduce-scatter, the exchange itself: merged[k - bucket_base] = sum over the nsrc sources (every GPU's chunks;
// srcs[0] is this GPU's own) of bucket k, for the PEER_TPB consecutive buckets of this block.  The PEER_TPB records of one
// source are contiguous (28 KB) whenever every bucket of the range is one piece -- always with uniform scalars -- and are
// pulled from the peer's HBM over NVLink by ONE TMA bulk copy (cp.async.bulk global -> shared, completion on an mbarrier)
// into a double buffer: the copy of source s + 1 is in flight while the threads add source s.  What that buys over SM-issued
// loads is the request size on the link (profiles/r02_peer_reduce.txt, 8 GPUs, 0.85 GB per rank): 16-byte loads per thread
// 7.5 ms for the stage, coalesced 16-byte cp.async 4.9 ms for this kernel (175 GB/s), bulk copies: see there.
// Buckets split into several pieces (skewed scalar sets) take the per-thread path.
#define PEER_TPB 128
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <class C>
__global__ void __launch_bounds__(PEER_TPB, 3) k_peer_merge(const red_src_t *__restrict__ srcs, uint32_t nsrc, uint32_t bucket_base, uint32_t nslice,
                                                            xyzz_t *__restrict__ merged, const uint32_t *__restrict__ wait_flags, uint32_t wait_n, uint32_t epoch) {
    extern __shared__ __align__(128) uint8_t sm_peer[];      // [2][PEER_TPB] records of 224 bytes, as they lie in the source
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t fast[2];
    if (threadIdx.x < wait_n)                                // the peers' bucket sums of this epoch must be complete
        while ((int32_t)(ld_acquire_sys(wait_flags + threadIdx.x) - epoch) < 0) __nanosleep(100);
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t b0 = blockIdx.x * PEER_TPB;               // first bucket of the block within the slice
    const uint32_t cnt = nslice - b0 < PEER_TPB ? nslice - b0 : PEER_TPB;
    const uint32_t k0 = bucket_base + b0, k = k0 + threadIdx.x;
    // thread 0 starts the copy of source s into buffer s & 1 (the buffer's previous readers have passed a __syncthreads)
    auto issue = [&](uint32_t s) {
        const uint32_t buf = s & 1;
        const uint32_t pf = __ldcv(srcs[s].poff + k0), pl = __ldcv(srcs[s].poff + k0 + cnt);
        const bool f = pl - pf == cnt;                       // every bucket has at least one piece: cnt pieces <=> one each
        fast[buf] = f ? 1u : 0u;
        if (f) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&bar[buf], cnt * (uint32_t)sizeof(xyzz_t));
            bulk_g2s(sm_peer + (size_t)buf * PEER_TPB * sizeof(xyzz_t), srcs[s].res + pf, cnt * (uint32_t)sizeof(xyzz_t), &bar[buf]);
        } else {
            mbar_arrive(&bar[buf]);                          // nothing to wait for: the phase completes at once
        }
    };
    xyzz_t acc, b;
    Grp<C>::set_inf(acc);
    if (threadIdx.x == 0) issue(0);
    for (uint32_t s = 0; s < nsrc; s++) {
        if (threadIdx.x == 0 && s + 1 < nsrc) issue(s + 1);
        const uint32_t buf = s & 1;
        mbar_wait(&bar[buf], (s >> 1) & 1);                  // the (s / 2)-th use of this buffer
        __syncthreads();                                     // fast[buf] written by thread 0 before its arrive
        if (threadIdx.x < cnt) {
            if (fast[buf]) {
                load_xyzz(b, reinterpret_cast<const xyzz_t *>(sm_peer + (size_t)buf * PEER_TPB * sizeof(xyzz_t)) + threadIdx.x);
                if (s == 0) acc = b; else Grp<C>::add_ni(acc, b);      // the first source is this GPU's own sum: nothing to add it to yet
            } else {
                const uint32_t p0 = __ldcv(srcs[s].poff + k), p1 = __ldcv(srcs[s].poff + k + 1);
                for (uint32_t q = p0; q < p1; q++) { load_xyzz_src(b, srcs[s].res + q, true); Grp<C>::add_ni(acc, b); }
            }
        }
        __syncthreads();                                     // the buffer is refilled two sources later
    }
    if (threadIdx.x < cnt) store_xyzz(merged + b0 + threadIdx.x, acc);
}
// out[g] = sum of in[g*group .. (g+1)*group)   (one block per group)
template <class C>
__global__ void __launch_bounds__(128) k_block_sum(const xyzz_t *__restrict__ in, uint32_t group, xyzz_t *__restrict__ out) {
    __shared__ xyzz_t sh[64];
    xyzz_t v, b;
    Grp<C>::set_inf(v);
    for (uint32_t k = threadIdx.x; k < group; k += blockDim.x) { load_xyzz(b, in + (size_t)blockIdx.x * group + k); Grp<C>::add_ni(v, b); }
    block_tree_sum<C>(v, sh);
    if (threadIdx.x == 0) store_xyzz(out + blockIdx.x, v);
}

// ---------------------------------------------------------------------------------------------
// a10: finalize (one block, one warp per term).  Per folded window f the host hands a list of terms, each an array of
// <= 32 points:   kind 0: 2^shift * sum(arr)        kind 1: 2^shift * sum_k k * arr[k]
// window value = sum of its terms; result = Horner over windows with c doublings; then either the XYZZ value is
// stored (partial result of one GPU) or it is exported as an arkworks G1Projective (export_jacobian).
#define FIN_WARPS 12     // one warp per term for the plans in use (8 - 11 terms): the longest term, not a warp's queue, is the critical path (static shared memory: 46 KB)
template <class C> __device__ void warp_sum(xyzz_t &v, xyzz_t *sh) {
    // tree sum of the 32 lanes' points through shared memory; result in lane 0
    const int lane = threadIdx.x & 31;
    for (int o = 16; o > 0; o >>= 1) {
        __syncwarp();
        if (lane >= o && lane < 2 * o) store_xyzz(sh + lane - o, v);
        __syncwarp();
        if (lane < o) { xyzz_t t; load_xyzz(t, sh + lane); Grp<C>::add_ni(v, t); }
    }
    __syncwarp();
}
// XYZZ -> arkworks G1Projective (Jacobian x, y, z; 3 x 6 x u64 Montgomery R = 2^384): the canonical affine point with
// z = 1, or the arkworks zero (1, 1, 0).  One inversion (~0.1 ms on one thread) buys a byte-deterministic result: the
// XYZZ representative depends on the order the scatter's atomics filled the buckets, the normalised point does not.
template <class C> __device__ void export_jacobian(const xyzz_t &total, uint32_t *out_jac) {
    fp x, y, z;
    Fp<C>::set_one(z);
    if (!Grp<C>::to_sw_affine(x, y, total)) { Fp<C>::set_one(x); Fp<C>::set_one(y); Fp<C>::set_zero(z); }
    fp_abi ox, oy, oz;
    Ec<C>::to_abi_ni(ox, x); Ec<C>::to_abi_ni(oy, y); Ec<C>::to_abi_ni(oz, z);
#pragma unroll
    for (int i = 0; i < 12; i++) { out_jac[i] = ox.w[i]; out_jac[12 + i] = oy.w[i]; out_jac[24 + i] = oz.w[i]; }
}
// One block per folded window f: the window's value (its terms, one warp each), times 2^(c f) by lane 0 (c f doublings: the
// serial part, so the windows do theirs side by side instead of one Horner chain of c (F - 1) doublings plus F term rounds in
// one block -- 9.4 ms of a 132 ms one-call MSM at 2^24 with F = 15, profiles/r01_one_call_launches_2pow24.txt); the block that
// finishes last sums the F products and exports.  F = 1 (every resident-bases plan that fits its tables): one block, no hand-over.
template <class C>
__global__ void __launch_bounds__(32 * FIN_WARPS) k_finalize(fin_args_t a, xyzz_t *__restrict__ out_xyzz, uint32_t *__restrict__ out_jac,
                                                           xyzz_t *__restrict__ scratch, uint32_t *__restrict__ ticket) {
    __shared__ xyzz_t sh[FIN_WARPS][16];
    __shared__ xyzz_t slot[FIN_WARPS];
    __shared__ uint32_t is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x;
    xyzz_t total;                       // meaningful in thread 0
    Grp<C>::set_inf(total);
    {
        xyzz_t part;                    // this warp's share of the window value (lane 0)
        Grp<C>::set_inf(part);
        for (int ti = warp; ti < a.nterms; ti += FIN_WARPS) {
            const fin_term_t &tm = a.term[ti];
            xyzz_t v;
            Grp<C>::set_inf(v);
            if ((uint32_t)lane < tm.n) {
                load_xyzz(v, tm.arr + (size_t)f * tm.stride_f + lane);
                if (tm.kind == 1) {         // (base + lane) * v by double-and-add over 5 bits (base + n <= 32)
                    xyzz_t r;
                    Grp<C>::set_inf(r);
                    const uint32_t wgt = tm.base + lane;
                    for (int b = 4; b >= 0; b--) { Grp<C>::dbl_ni(r); if ((wgt >> b) & 1) Grp<C>::add_ni(r, v); }
                    v = r;
                }
            }
            warp_sum<C>(v, sh[warp]);
            if (lane == 0) {
                for (uint32_t s = 0; s < tm.shift; s++) Grp<C>::dbl_ni(v);
                Grp<C>::add_ni(part, v);
            }
        }
        if (lane == 0) store_xyzz(&slot[warp], part);
        __syncthreads();
        if (warp == 0) {
            xyzz_t v;
            Grp<C>::set_inf(v);
            if (lane < FIN_WARPS) load_xyzz(v, &slot[lane]);
            warp_sum<C>(v, sh[0]);
            if (lane == 0) total = v;
        }
    }
    if (gridDim.x > 1) {
        if (threadIdx.x == 0) {
            for (int s = 0; s < a.c * f; s++) Grp<C>::dbl_ni(total);
            store_xyzz(scratch + f, total);
            __threadfence();
            is_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
        }
        __syncthreads();
        if (!is_last || warp != 0) return;
        __threadfence();
        xyzz_t v;
        Grp<C>::set_inf(v);
        for (uint32_t k = lane; k < gridDim.x; k += 32) {
            xyzz_t w;
            const uint4 *src = reinterpret_cast<const uint4 *>(scratch + k);
            uint4 *t = reinterpret_cast<uint4 *>(&w);
#pragma unroll
            for (int i = 0; i < XYZZ_VEC; i++) t[i] = __ldcg(src + i);
            Grp<C>::add_ni(v, w);
        }
        warp_sum<C>(v, sh[0]);
        if (lane == 0) { total = v; *ticket = 0; }
    }
    if (warp != 0 || lane != 0) return;
    if (out_xyzz) store_xyzz(out_xyzz, total);
    if (out_jac) export_jacobian<C>(total, out_jac);
}
// a10 (multi-GPU): sum `count` partial results (one XYZZ record per GPU, gathered over NVLink) and export.  One warp.
template <class C>
__global__ void __launch_bounds__(32) k_combine(const xyzz_t *__restrict__ parts, int count, uint32_t *__restrict__ out_jac) {
    __shared__ xyzz_t sh[16];
    const int lane = threadIdx.x;
    xyzz_t v;
    Grp<C>::set_inf(v);
    for (int k = lane; k < count; k += 32) { xyzz_t t; load_xyzz(t, parts + k); Grp<C>::add_ni(v, t); }
    warp_sum<C>(v, sh);
    if (lane == 0) export_jacobian<C>(v, out_jac);
}

// ---------------------------------------------------------------------------------------------
// a11: base tables.
// arkworks G1Affine records (x, y: 6 x u64 Montgomery R = 2^384; infinity flag at byte 96; stride bytes) -> table 0 in
// the device representation (14 x 28-bit limbs, R = 2^392) + infinity bitmask
template <class C>
__global__ void __launch_bounds__(128) k_convert_points(const uint8_t *__restrict__ src, uint32_t n, uint32_t stride, affine_t *__restrict__ table0, uint32_t *__restrict__ infmask) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool inf = false;
    if (i < n) {
        const uint32_t *p = reinterpret_cast<const uint32_t *>(src + (size_t)i * stride);   // 104 and 96 are multiples of 8
        fp_abi ax, ay;
#pragma unroll
        for (int k = 0; k < 12; k++) { ax.w[k] = p[k]; ay.w[k] = p[12 + k]; }
        affine_t o;
        Ec<C>::from_abi_ni(o.x, ax);
        Ec<C>::from_abi_ni(o.y, ay);
        store_affine(table0 + i, o);
        // infinity: the arkworks flag byte (records longer than 96 bytes), or x = y = 0 -- the encoding of the flag-less
        // 96-byte blst_p1_affine records (sppark affine_t::is_inf, P1A/6block/sppark/ec/xyzz_t.hpp:44-45); (0, 0) is on
        // neither curve, so the second test never misreads a finite point
        uint32_t any = 0;
#pragma unroll
        for (int k = 0; k < 24; k++) any |= p[k];
        inf = (stride > 96 && src[(size_t)i * stride + 96] != 0) || any == 0;
    }
    uint32_t bits = __ballot_sync(0xffffffffu, inf);
    if ((threadIdx.x & 31) == 0 && i < n) infmask[i >> 5] = bits;
}
// table 0 back as arkworks (x, y) records of 96 bytes (tests)
template <class C>
__global__ void __launch_bounds__(128) k_export_points(const typename Grp<C>::entry_t *__restrict__ entries, uint32_t n, uint32_t *__restrict__ dst, bool ark) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp x, y;
    Grp<C>::entry_to_sw(x, y, entries[i], ark);
    fp_abi ox, oy;
    Ec<C>::to_abi_ni(ox, x);
    Ec<C>::to_abi_ni(oy, y);
#pragma unroll
    for (int k = 0; k < 12; k++) { dst[(size_t)i * 24 + k] = ox.w[k]; dst[(size_t)i * 24 + 12 + k] = oy.w[k]; }
}
// short-Weierstrass affine chunk -> table entries in the curve's bucket coordinates (grp.cuh)
template <class C>
__global__ void __launch_bounds__(128) k_make_entries(const affine_t *__restrict__ src, uint32_t n, typename Grp<C>::entry_t *__restrict__ dst, bool ark,
                                                      const uint32_t *__restrict__ infmask, uint32_t *__restrict__ bad) {
    Grp<C>::make_entries(src, n, dst, blockIdx.x * blockDim.x + threadIdx.x, ark, infmask, bad);
}
// Opt-in subgroup check (b200msm_check_bases_in_g1): r * P_i for every base through the bucket arithmetic's own doubling and
// mixed add on table 0 (253 / 255 doublings + popcount(r) mixed adds per point: seconds at 2^26, init time only); counts the
// bases whose multiple is not the neutral element.  arkworks makes the same check when it deserialises a G1Affine
// (P1B/nickray/driver/algebra/ec/src/models/short_weierstrass_jacobian.rs is_in_correct_subgroup_assuming_on_curve).
template <class C>
__global__ void __launch_bounds__(128) k_check_order_r(const typename Grp<C>::entry_t *__restrict__ entries, const uint32_t *__restrict__ infmask, uint32_t n, bool ark,
                                                       uint32_t *__restrict__ outside) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || ((infmask[i >> 5] >> (i & 31)) & 1)) return;
    xyzz_t acc;
    Grp<C>::set_inf(acc);
    for (int b = C::SCALAR_BITS - 1; b >= 0; b--) {
        Grp<C>::dbl_ni(acc);
        uint32_t word = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) if (w == (b >> 5)) word = C::r(w);
        if ((word >> (b & 31)) & 1) Grp<C>::madd_entry_ni(acc, entries + i, ark);
    }
    if (!Grp<C>::is_neutral(acc)) atomicAdd(outside, 1u);
}
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
// A ZZZ of zero -- 2^k P at infinity: P is a base flagged infinity, or a base outside the prime-order subgroup whose order
// divides 2^k -- is replaced by 1 so it cannot poison its batch, and the point is written as (0, 0): the XYZZ mixed add
// skips such entries (grp.cuh GrpXyzz::madd_staged), so that arithmetic stays exact for every on-curve base; the
// twisted-Edwards arithmetic has no such entry and counts the base in *bad instead (unless infmask flags it).
template <class C>
__global__ void __launch_bounds__(128) k_batch_normalize(const xyzz_t *__restrict__ tmp, uint32_t n, affine_t *__restrict__ dst,
                                                         const uint32_t *__restrict__ infmask, uint32_t *__restrict__ bad) {
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
        const uint2 *s = reinterpret_cast<const uint2 *>(&tmp[base + k].zzz);     // 56-byte field: 8-byte aligned
        uint2 *d = reinterpret_cast<uint2 *>(&z);
#pragma unroll
        for (int v = 0; v < 7; v++) d[v] = s[v];
        if (F::is_zero(z)) {
            F::set_one(z);
            const uint32_t i = base + k;
            if (!Grp<C>::INF_ENTRIES && bad && !(infmask && (infmask[i >> 5] >> (i & 31)) & 1)) atomicAdd(bad, 1u);
        }
        prefix[k] = acc;
        Ec<C>::mul_ni(acc, acc, z);
    }
    fp inv;
    Ec<C>::inv_ni(inv, acc);
    for (uint32_t k = cnt; k-- > 0;) {
        xyzz_t q;
        load_xyzz(q, tmp + base + k);
        const bool at_inf = F::is_zero(q.zzz);
        if (at_inf) F::set_one(q.zzz);
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
        F::csub_p(o.x);
        F::csub_p(o.y);
        if (at_inf) { F::set_zero(o.x); F::set_zero(o.y); }
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
// test entry points: element-wise field / curve operations.  Operands and results are arkworks-ABI field elements
// (48 bytes, R = 2^384, canonical); XYZZ records are 4 such elements (192 bytes), affine 2 (96 bytes).
template <class C>
__global__ void k_test_fp(int op, const fp_abi *a, const fp_abi *b, fp_abi *r, uint32_t n) {
    typedef Fp<C> F;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp x, y, t;
    Ec<C>::from_abi_ni(x, a[i]);
    Ec<C>::from_abi_ni(y, b[i]);
    switch (op) {
    case 0: F::mul(t, x, y); break;
    case 1: F::add(t, x, y); break;
    case 2: F::template sub<2>(t, x, y); break;
    case 3: F::sqr(t, x); break;
    case 4: Ec<C>::inv_ni(t, x); break;
    default: t = x;                       // 5: ABI round trip
    }
    Ec<C>::to_abi_ni(r[i], t);
}
template <class C>
__global__ void k_test_ec(int op, xyzz_abi_t *acc, const void *other, uint32_t n, int reps) {
    typedef Ec<C> E;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xyzz_t q, o;
    affine_t a;
    E::from_abi_ni(q.x, acc[i].x); E::from_abi_ni(q.y, acc[i].y); E::from_abi_ni(q.zz, acc[i].zz); E::from_abi_ni(q.zzz, acc[i].zzz);
    if (op == 0 || op == 3) { E::from_abi_ni(a.x, ((const affine_abi_t *)other)[i].x); E::from_abi_ni(a.y, ((const affine_abi_t *)other)[i].y); }
    if (op == 1) {
        const xyzz_abi_t &s = ((const xyzz_abi_t *)other)[i];
        E::from_abi_ni(o.x, s.x); E::from_abi_ni(o.y, s.y); E::from_abi_ni(o.zz, s.zz); E::from_abi_ni(o.zzz, s.zzz);
    }
    for (int k = 0; k < reps; k++) {
        if (op == 0) E::madd(q, a.x, a.y);          // the inlined form the accumulate kernel uses
        else if (op == 1) E::add_ni(q, o);
        else if (op == 2) E::dbl_ni(q);
        else if (op == 3) E::dbl_affine_ni(q, a.x, a.y);
    }
    E::to_abi_ni(acc[i].x, q.x); E::to_abi_ni(acc[i].y, q.y); E::to_abi_ni(acc[i].zz, q.zz); E::to_abi_ni(acc[i].zzz, q.zzz);
}

// group-law test: a, b short-Weierstrass affine (arkworks 96-byte records) -> entry / bucket coordinates -> op ->
// back to an arkworks G1Affine record (104 bytes: x, y, infinity flag).  op 0: a + reps x b through the mixed add on
// a table entry; 1: the same through the full add; 2: 2^reps a; 3: a - b (negated entry).
template <class C>
__global__ void k_test_grp(int op, const affine_abi_t *a, const affine_abi_t *b, uint32_t *out104, uint32_t n, int reps) {
    typedef Grp<C> G;
    __shared__ affine_t sw[2][64];
    __shared__ typename G::entry_t ent[2][64];
    __shared__ uint4 stage[G::ENTRY_VEC][64];
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x;
    const uint32_t srci = i < n ? i : 0;         // pad the conversion batch with a valid point
    Ec<C>::from_abi_ni(sw[0][lane].x, a[srci].x); Ec<C>::from_abi_ni(sw[0][lane].y, a[srci].y);
    Ec<C>::from_abi_ni(sw[1][lane].x, b[srci].x); Ec<C>::from_abi_ni(sw[1][lane].y, b[srci].y);
    __syncthreads();
    // GRP_BATCH consecutive points per converting thread
    if (lane < 64 / GRP_BATCH) { G::make_entries(sw[0], 64, ent[0], lane, false, nullptr, nullptr); G::make_entries(sw[1], 64, ent[1], lane, false, nullptr, nullptr); }
    __syncthreads();
    if (i >= n) return;
    xyzz_t acc, other;
    G::set_inf(acc);
    G::set_inf(other);
    auto put = [&](const typename G::entry_t &e) {
        const uint4 *s = reinterpret_cast<const uint4 *>(&e);
#pragma unroll
        for (int k = 0; k < G::ENTRY_VEC; k++) stage[k][lane] = s[k];
    };
    put(ent[0][lane]);
    G::madd_staged(acc, &stage[0][lane], 64, false);             // acc = a
    put(ent[1][lane]);
    G::madd_staged(other, &stage[0][lane], 64, false);           // other = b
    for (int k = 0; k < reps; k++) {
        if (op == 0) G::madd_staged(acc, &stage[0][lane], 64, false);
        else if (op == 1) G::add_ni(acc, other);
        else if (op == 2) G::dbl_ni(acc);
        else G::madd_staged(acc, &stage[0][lane], 64, true);
    }
    fp x, y;
    const bool finite = G::to_sw_affine(x, y, acc);
    fp_abi ox, oy;
    if (!finite) { Fp<C>::set_zero(x); Fp<C>::set_one(y); }
    Ec<C>::to_abi_ni(ox, x); Ec<C>::to_abi_ni(oy, y);
    uint32_t *o = out104 + (size_t)i * 26;
#pragma unroll
    for (int k = 0; k < 12; k++) { o[k] = ox.w[k]; o[12 + k] = oy.w[k]; }
    o[24] = finite ? 0u : 1u;
    o[25] = 0;
}

}  // namespace b200msm

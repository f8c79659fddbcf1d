// kernels_hot.cuh -- the per-MSM kernels that touch every scalar / every point (templated on the curve):
//   k_digits<PASS>   a4  signed-digit window decomposition fused with the bucket histogram (PASS 0) and with the
//                        scatter of point references into per-bucket lists (PASS 1)
//   k_accumulate     a6  the hot loop: thread = piece of a bucket, XYZZ += affine, points prefetched with cp.async
// (a4 / a6 = rows of SURVEY.md section 8a; reference counterparts are listed in DESIGN.md.)
#pragma once
#include "kernels_util.cuh"
#include "launch.h"
#include "digits.cuh"

namespace b200msm {

// ---------------------------------------------------------------------------------------------
// a4: scalar -> signed window digits.
//
//   s <- s mod r;  if the top scalar bit is set, s <- r - s and every digit's sign flips ((r-s)(-P) = sP), so
//   s < 2^(SCALAR_BITS-1) and W = ceil(SCALAR_BITS / c) windows absorb the last carry.
//   digit d_w in [-2^(c-1), 2^(c-1)]:  raw + carry > 2^(c-1)  ->  d = raw + carry - 2^c, carry 1.
//   window w is served from base table j = w / F (holding 2^(c*F*j) P) into folded window f = w % F:
//   bucket key = f * 2^(c-1) + |d| - 1,  point reference = j*N + i, bit 31 = negate.
// PASS 0 counts keys; PASS 1 takes a slot in the bucket's list (cursor pre-loaded with the bucket offsets).
template <class C, int PASS>
__global__ void __launch_bounds__(256) k_digits(const uint32_t *__restrict__ scalars, const uint32_t *__restrict__ infmask,
                                                uint32_t n, uint32_t first, uint32_t table_stride, int c, int W, int F,
                                                uint32_t *__restrict__ counters, uint32_t *__restrict__ refs) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;       // scalars[i] belongs to base first + i (chunked calls)
    if (i >= n) return;
    const uint32_t gi = first + i;
    if (infmask[gi >> 5] & (1u << (gi & 31))) return;     // base at infinity contributes nothing
    if (PASS == 0 && W <= 16) {        // large problems: compile-time trip count
        // A warp whose lanes all hit the SAME counter (skewed scalar sets: all-equal, tiny, sparse -- SURVEY A3) adds once for the
        // warp: 2^24 equal scalars otherwise serialise 1.8e8 atomics on 11 addresses (digits_hist 14.8 ms, profiles/r02_skew.txt)
        for_each_digit<C, 16>(scalars, i, gi, table_stride, c, W, F, [&](int, uint32_t key, uint32_t) {
            const uint32_t act = __activemask();
            const uint32_t k0 = __shfl_sync(act, key, __ffs(act) - 1);
            if (__all_sync(act, key == k0)) { if ((threadIdx.x & 31) == (uint32_t)(__ffs(act) - 1)) atomicAdd(&counters[key], (uint32_t)__popc(act)); }
            else atomicAdd(&counters[key], 1u);
        });
        return;
    }
    for_each_digit<C>(scalars, i, gi, table_stride, c, W, F, [&](uint32_t key, uint32_t ref) {
        if (PASS == 0) {
            atomicAdd(&counters[key], 1u);
        } else {
            uint32_t slot = atomicAdd(&counters[key], 1u);
            refs[slot] = ref;
        }
    });
}

// Two-level scatter for large problems (a5; stands in for yrrid's partition1024 / partition4096 pair,
// Partition1024.cu:157-243, Partition4096.cu:352-433).  A direct scatter of W*n references over a 3 GB list array is
// one 32-byte sector read-modify-write in DRAM per 4-byte reference.  Instead:
//   k_digits_bin   groups (key, ref) pairs by bin = key >> bin_shift (<= 256 bins, each bin's share of the list array
//                  sized to stay L2-resident): per block, count per bin in shared memory, reserve the block's range in
//                  every bin with one global atomic per bin, then write the pairs (8 bytes each, ~W*256/bins
//                  consecutive pairs per bin per block);
//   k_bin_scatter  streams the pairs in order and does the atomic-slot scatter; all blocks in flight work inside one or
//                  two bins, so cursors and list sectors are hit in L2 and reach DRAM once, as full sectors.
#define BIN_MAX 256
#define BIN_TPB 256
#define BIN_ITEMS_MAX (BIN_TPB * 16)      // the staged path needs W <= 16 (it is only used for large n, where c >= 16)
template <class C>
__global__ void __launch_bounds__(BIN_TPB) k_digits_bin(const uint32_t *__restrict__ scalars, const uint32_t *__restrict__ infmask,
                                                        uint32_t n, uint32_t first, uint32_t table_stride, int c, int W, int F, int bin_shift, uint32_t nbins,
                                                        uint32_t *__restrict__ bin_cursor, uint2 *__restrict__ staging) {
    __shared__ uint32_t cnt[BIN_MAX], start[BIN_MAX], base[BIN_MAX];
    __shared__ uint32_t scan[BIN_TPB];
    __shared__ uint2 sorted[BIN_ITEMS_MAX];                     // this block's pairs grouped by bin (32 KB)
    for (uint32_t b = threadIdx.x; b < BIN_MAX; b += BIN_TPB) cnt[b] = 0;
    __syncthreads();
    uint32_t i = blockIdx.x * BIN_TPB + threadIdx.x;
    const uint32_t gi = first + i;
    const bool live = i < n && !(infmask[gi >> 5] & (1u << (gi & 31)));
    // the digits are computed once and kept in registers for both passes (W <= 16 on this path)
    uint32_t key[16], ref[16];
#pragma unroll
    for (int w = 0; w < 16; w++) key[w] = 0xffffffffu;
    if (live) for_each_digit<C, 16>(scalars, i, gi, table_stride, c, W, F, [&](int w, uint32_t k, uint32_t r) { key[w] = k; ref[w] = r; });
#pragma unroll
    for (int w = 0; w < 16; w++) if (key[w] != 0xffffffffu) atomicAdd(&cnt[key[w] >> bin_shift], 1u);
    __syncthreads();
    // exclusive scan of the per-bin counts (BIN_MAX == BIN_TPB: one bin per thread) and one global reservation per bin
    {
        const uint32_t k = cnt[threadIdx.x];
        scan[threadIdx.x] = k;
        __syncthreads();
        for (int o = 1; o < BIN_TPB; o <<= 1) {
            uint32_t t = threadIdx.x >= (uint32_t)o ? scan[threadIdx.x - o] : 0;
            __syncthreads();
            scan[threadIdx.x] += t;
            __syncthreads();
        }
        start[threadIdx.x] = scan[threadIdx.x] - k;
        base[threadIdx.x] = (k && threadIdx.x < nbins) ? atomicAdd(&bin_cursor[threadIdx.x], k) : 0;
        cnt[threadIdx.x] = 0;
    }
    __syncthreads();
    const uint32_t total = scan[BIN_TPB - 1];
#pragma unroll
    for (int w = 0; w < 16; w++) if (key[w] != 0xffffffffu) {
        const uint32_t b = key[w] >> bin_shift;
        sorted[start[b] + atomicAdd(&cnt[b], 1u)] = make_uint2(key[w], ref[w]);
    }
    __syncthreads();
    // consecutive threads write consecutive pairs of a bin's run: ~W*256/bins pairs (x 8 bytes) per run
    for (uint32_t k = threadIdx.x; k < total; k += BIN_TPB) {
        const uint2 v = sorted[k];
        const uint32_t b = v.x >> bin_shift;
        staging[base[b] + (k - start[b])] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// a6: bucket accumulation.  One thread sums one piece: acc (registers) += table[ref] for each reference in the piece,
// in the curve's bucket coordinates (grp.cuh: twisted-Edwards 7 M unified add on BLS12-377, XYZZ on BLS12-381).
// The next entry (144 / 112 bytes) is fetched with 16-byte cp.async into the thread's own shared-memory slots
// ([stage][16-byte chunk][thread]: conflict-free) while the current mixed add (~5000 integer instructions) runs, so the
// random gather latency is hidden by one stage of prefetch.  Results go to res[piece].
#if defined(ACC_TMA)
// Experiment switch (-DACC_TMA=1; north_star asks for TMA staging, DESIGN.md section 6 explains why the default is cp.async):
// the entry gather as one TMA bulk copy per lane (cp.async.bulk global -> shared, 144 contiguous bytes into the thread's own
// slot, completion counted on a per-warp mbarrier per stage), two stages.  Same results; measured slower than the nine 16-byte
// cp.async per lane: UBLKCP takes uniform-register operands, so a warp's 32 different source addresses are issued lane by
// lane (profiles/r02_acc_tma.txt has the SASS and the timing).
__device__ __forceinline__ uint32_t hot_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <class C, class ACC>
__device__ __forceinline__ void accumulate_piece_tma(const typename Grp<C>::entry_t *__restrict__ tables, const uint32_t *__restrict__ refs,
                                                     uint32_t start, uint32_t len, uint4 *slots /* [2][ACC_TPB][EV] */, uint64_t *bars /* [2][warps] */, xyzz_t *__restrict__ dst) {
    constexpr int EV = Grp<C>::ENTRY_VEC;
    constexpr uint32_t BYTES = EV * 16;
    const uint32_t warp = threadIdx.x >> 5;
    typename ACC::acc_t acc;
    ACC::init(acc);
    // every lane of the warp walks the same number of iterations (the longest piece of the warp): the barrier counts 32 arrivals
    uint32_t wlen = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const uint32_t t = __shfl_xor_sync(0xffffffffu, wlen, o); wlen = t > wlen ? t : wlen; }
    auto issue = [&](uint32_t i) {
        const uint32_t st = i & 1;
        uint64_t *bar = bars + st * (ACC_TPB / 32) + warp;
        if (i < len) {
            const uint32_t ref = __ldg(refs + start + i);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the slot's last generic reads precede the asynchronous refill
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hot_smem_u32(bar)), "r"(BYTES) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(hot_smem_u32(slots + ((size_t)st * ACC_TPB + threadIdx.x) * EV)),
                         "l"(tables + (ref & 0x7fffffffu)), "r"(BYTES), "r"(hot_smem_u32(bar)) : "memory");
            return ref;
        }
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(hot_smem_u32(bar)) : "memory");
        return 0u;
    };
    uint32_t ref = issue(0), ref_next = 0;
    for (uint32_t i = 0; i < wlen; i++) {
        if (i + 1 < wlen) ref_next = issue(i + 1);
        uint64_t *bar = bars + (i & 1) * (ACC_TPB / 32) + warp;
        uint32_t done;
        do {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(hot_smem_u32(bar)), "r"((i >> 1) & 1) : "memory");
        } while (!done);
        if (i < len) ACC::madd_staged(acc, slots + ((size_t)(i & 1) * ACC_TPB + threadIdx.x) * EV, 1, (ref >> 31) != 0);
        __syncwarp();                        // every lane has read its slot before the stage is refilled
        ref = ref_next;
    }
    if (!dst) return;
    xyzz_t out;
    ACC::export_to(out, acc);
    store_xyzz(dst, out);
}
#endif
template <class C, class ACC>
__device__ __forceinline__ void accumulate_piece(const typename Grp<C>::entry_t *__restrict__ tables, const uint32_t *__restrict__ refs,
                                                 uint32_t start, uint32_t len, uint4 (*stage)[Grp<C>::ENTRY_VEC][ACC_TPB], xyzz_t *__restrict__ dst) {
    constexpr int EV = Grp<C>::ENTRY_VEC;
    typename ACC::acc_t acc;
    ACC::init(acc);
    // The list of references is read TWO iterations ahead: the reference of entry i + 1 is already in a register when its
    // gather is issued (loading it there stalled the warp for an L2 / DRAM round trip per mixed add: long_scoreboard 0.28 per
    // issue in profiles/r02_ncu_accumulate_fp29.txt).
    uint32_t ref = 0, ref_next = 0, ref_next2 = 0;
    if (len) {
        ref = __ldg(refs + start);
        if (len > 1) ref_next = __ldg(refs + start + 1);
        const uint4 *src = reinterpret_cast<const uint4 *>(tables + (ref & 0x7fffffffu));
#pragma unroll
        for (int k = 0; k < EV; k++) cp_async16(&stage[0][k][threadIdx.x], src + k);
    }
    cp_async_commit();
    for (uint32_t i = 0; i < len; i++) {
        if (i + 1 < len) {
            const uint4 *src = reinterpret_cast<const uint4 *>(tables + (ref_next & 0x7fffffffu));
#pragma unroll
            for (int k = 0; k < EV; k++) cp_async16(&stage[(i + 1) & 1][k][threadIdx.x], src + k);
        }
        if (i + 2 < len) ref_next2 = __ldg(refs + start + i + 2);
        cp_async_commit();
        cp_async_wait<1>();
        ACC::madd_staged(acc, &stage[i & 1][0][threadIdx.x], ACC_TPB, (ref >> 31) != 0);
        ref = ref_next;
        ref_next = ref_next2;
    }
    xyzz_t out;
    ACC::export_to(out, acc);
    store_xyzz(dst, out);
}
// FP64 = false: the group's integer arithmetic (default).  FP64 = true (curves with a twisted-Edwards model): the same
// loop with the accumulator on the FP64 pipe (te.cuh TeD) -- measured equal to the integer flavour in this kernel
// (both ~96 SM-clocks per mixed add at 2^26) and the two do not overlap when mixed on one SM, so
// it is kept as an experiment switch (B200MSM_ACC=fp64), not the default.
template <class C, bool FP64>
__global__ void __launch_bounds__(ACC_TPB, ACC_MINB) k_accumulate(const typename Grp<C>::entry_t *__restrict__ tables, const uint32_t *__restrict__ refs,
                                                           const uint32_t *__restrict__ order, const uint32_t *__restrict__ piece_start,
                                                           const uint32_t *__restrict__ piece_len, const uint32_t *__restrict__ npieces_ptr,
                                                           xyzz_t *__restrict__ res) {
    typedef Grp<C> G;
    uint32_t j = blockIdx.x * ACC_TPB + threadIdx.x;
#if defined(ACC_TMA)
    if constexpr (C::HAS_TE && !FP64) {
        __shared__ __align__(128) uint4 slots[2 * ACC_TPB * G::ENTRY_VEC];
        __shared__ __align__(8) uint64_t bars[2 * (ACC_TPB / 32)];
        if (threadIdx.x < 2 * (ACC_TPB / 32)) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hot_smem_u32(&bars[threadIdx.x])), "r"(32) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
        const bool live = j < *npieces_ptr;                    // whole warps stay: the per-warp barriers count 32 arrivals
        const uint32_t piece = live ? order[j] : 0;
        accumulate_piece_tma<C, AccTe29<C>>(tables, refs, live ? piece_start[piece] : 0, live ? piece_len[piece] : 0, slots, bars, live ? res + piece : nullptr);
        return;
    }
#endif
    __shared__ uint4 stage[2][G::ENTRY_VEC][ACC_TPB];
    if (j >= *npieces_ptr) return;
    const uint32_t piece = order[j];
    const uint32_t start = piece_start[piece], len = piece_len[piece];
    if constexpr (FP64 && C::HAS_TE) accumulate_piece<C, AccFp64Te<C>>(tables, refs, start, len, stage, res + piece);
    else if constexpr (C::HAS_TE) accumulate_piece<C, AccTe29<C>>(tables, refs, start, len, stage, res + piece);      // 13 x 29-bit field (fp29.cuh)
    else accumulate_piece<C, AccNative<G>>(tables, refs, start, len, stage, res + piece);
}

}  // namespace b200msm

// kernels_common.cuh -- curve-independent kernels (included by msm.cu only):
//   k_scan_*         a5  exclusive prefix sums (bucket offsets, piece offsets)
//   k_piece_*        a5  split buckets into pieces of at most Lmax points, order pieces by decreasing length
//   k_replicate      a11 test bases: copy the first entries of table 0 up to n
#pragma once
#include "kernels_util.cuh"

namespace b200msm {

__device__ __forceinline__ uint32_t ld_acquire_sys_c(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_c(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// exclusive prefix sum of uint32 (three launches; blocks of 256 threads x 8 elements)
#define SCAN_TPB 256
#define SCAN_EPT 8
template <int MODE>   // MODE 0: in = raw counts;  MODE 1: in = pieces per bucket = max(1, ceil(count / lmax))
__device__ __forceinline__ uint32_t scan_input(uint32_t v, uint32_t lmax) {
    if (MODE == 0) return v;
    return v == 0 ? 1u : (v + lmax - 1) / lmax;
}
template <int MODE>
__global__ void __launch_bounds__(SCAN_TPB) k_scan_block_sums(const uint32_t *__restrict__ in, uint32_t n, uint32_t lmax, uint32_t *__restrict__ block_sums) {
    __shared__ uint32_t sh[SCAN_TPB / 32];
    uint32_t base = blockIdx.x * (SCAN_TPB * SCAN_EPT) + threadIdx.x * SCAN_EPT;
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_EPT; k++) if (base + k < n) sum += scan_input<MODE>(in[base + k], lmax);
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int k = 0; k < SCAN_TPB / 32; k++) t += sh[k];
        block_sums[blockIdx.x] = t;
    }
}
// single block: exclusive scan of block_sums in place; total written to block_sums[nblocks]
__global__ void __launch_bounds__(1024) k_scan_top(uint32_t *block_sums, uint32_t nblocks) {
    __shared__ uint32_t sh[1024];
    __shared__ uint32_t running;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nblocks; base += 1024) {
        uint32_t idx = base + threadIdx.x;
        uint32_t v = idx < nblocks ? block_sums[idx] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        uint32_t incl = sh[threadIdx.x];
        uint32_t r0 = running;
        if (idx < nblocks) block_sums[idx] = r0 + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) running = r0 + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[nblocks] = running;
}
// out[i] = exclusive prefix;  out[n] = total;  optional second copy (the scatter cursors)
template <int MODE>
__global__ void __launch_bounds__(SCAN_TPB) k_scan_apply(const uint32_t *__restrict__ in, uint32_t n, uint32_t lmax, const uint32_t *__restrict__ block_sums,
                                                        uint32_t *__restrict__ out, uint32_t *__restrict__ out2) {
    __shared__ uint32_t sh[SCAN_TPB];
    uint32_t base = blockIdx.x * (SCAN_TPB * SCAN_EPT) + threadIdx.x * SCAN_EPT;
    uint32_t v[SCAN_EPT], sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_EPT; k++) { v[k] = (base + k < n) ? scan_input<MODE>(in[base + k], lmax) : 0; sum += v[k]; }
    sh[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < SCAN_TPB; o <<= 1) {
        uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t run = block_sums[blockIdx.x] + sh[threadIdx.x] - sum;
#pragma unroll
    for (int k = 0; k < SCAN_EPT; k++) {
        if (base + k < n) { out[base + k] = run; if (out2) out2[base + k] = run; }
        run += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_TPB - 1) out[n] = block_sums[gridDim.x];
}

// ---------------------------------------------------------------------------------------------
// a5: pieces.  Bucket k owns pieces [piece_off[k], piece_off[k+1]); piece s of bucket k covers
// refs[bucket_off[k] + s*lmax ...] (at most lmax entries).  With uniform scalars every bucket is one piece; the
// split only exists so that a skewed scalar set (all equal, sparse, ...) cannot serialise on one thread.
// Pieces are then ordered by decreasing length (counting sort) so the 32 lanes of a warp run the same trip count.
#define PIECE_TPB 256
__global__ void __launch_bounds__(PIECE_TPB) k_piece_fill(const uint32_t *__restrict__ bucket_off, const uint32_t *__restrict__ piece_off, uint32_t nbuckets,
                                                          uint32_t lmax, uint32_t *__restrict__ piece_start, uint32_t *__restrict__ piece_len,
                                                          uint32_t *__restrict__ len_hist) {
    extern __shared__ uint32_t sh_hist[];   // lmax + 1 bins
    for (uint32_t b = threadIdx.x; b <= lmax; b += blockDim.x) sh_hist[b] = 0;
    __syncthreads();
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nbuckets) {
        uint32_t start = bucket_off[k], cnt = bucket_off[k + 1] - start;
        uint32_t j = piece_off[k], np = piece_off[k + 1] - j;
        for (uint32_t s = 0; s < np; s++) {
            uint32_t len = cnt > lmax ? lmax : cnt;
            piece_start[j + s] = start;
            piece_len[j + s] = len;
            atomicAdd(&sh_hist[len], 1u);
            start += len;
            cnt -= len;
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b <= lmax; b += blockDim.x) if (sh_hist[b]) atomicAdd(&len_hist[b], sh_hist[b]);
}
// single block: cursor[len] = number of pieces strictly longer than len (descending order start positions)
__global__ void __launch_bounds__(1024) k_piece_len_scan(const uint32_t *__restrict__ len_hist, uint32_t lmax, uint32_t *__restrict__ len_cursor) {
    __shared__ uint32_t sh[1024];
    __shared__ uint32_t running;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    // walk lengths from lmax down to 0 in tiles of 1024
    for (uint32_t base = 0; base <= lmax; base += 1024) {
        uint32_t r = base + threadIdx.x;                  // rank in descending order
        uint32_t v = r <= lmax ? len_hist[lmax - r] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        uint32_t r0 = running, incl = sh[threadIdx.x];
        if (r <= lmax) len_cursor[lmax - r] = r0 + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) running = r0 + incl;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(PIECE_TPB) k_piece_order(const uint32_t *__restrict__ piece_len, const uint32_t *__restrict__ npieces_ptr, uint32_t lmax,
                                                           uint32_t *__restrict__ len_cursor, uint32_t *__restrict__ order) {
    extern __shared__ uint32_t sh[];        // [0, lmax]: local histogram, then reserved global base per length
    for (uint32_t b = threadIdx.x; b <= lmax; b += blockDim.x) sh[b] = 0;
    __syncthreads();
    uint32_t np = *npieces_ptr;
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t len = 0, rank = 0;
    if (j < np) { len = piece_len[j]; rank = atomicAdd(&sh[len], 1u); }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b <= lmax; b += blockDim.x) { uint32_t cnt = sh[b]; if (cnt) sh[b] = atomicAdd(&len_cursor[b], cnt); }
    __syncthreads();
    if (j < np) order[sh[len] + rank] = j;
}

// ---------------------------------------------------------------------------------------------
// a5, two-level scatter (see kernels_hot.cuh k_digits_bin)
// bin b's range of the staging array starts where its first bucket's list starts
__global__ void __launch_bounds__(256) k_bin_init(const uint32_t *__restrict__ bucket_off, uint32_t nbuckets, int bin_shift, uint32_t nbins, uint32_t *__restrict__ bin_cursor) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nbins) {
        uint32_t first = b << bin_shift;
        bin_cursor[b] = bucket_off[first < nbuckets ? first : nbuckets];
    }
}
#define BINSC_TPB 256
#define BINSC_EPT 8
__global__ void __launch_bounds__(BINSC_TPB) k_bin_scatter(const uint2 *__restrict__ staging, const uint32_t *__restrict__ total_ptr,
                                                          uint32_t *__restrict__ cursor, uint32_t *__restrict__ refs) {
    const uint32_t total = *total_ptr;
    const size_t base = (size_t)blockIdx.x * (BINSC_TPB * BINSC_EPT);
    if (base >= total) return;
    uint2 v[BINSC_EPT];
#pragma unroll
    for (int k = 0; k < BINSC_EPT; k++) {
        size_t idx = base + (size_t)k * BINSC_TPB + threadIdx.x;          // coalesced 8-byte loads
        v[k] = idx < total ? __ldcs(staging + idx) : make_uint2(0xffffffffu, 0);
    }
    uint32_t slot[BINSC_EPT];
#pragma unroll
    for (int k = 0; k < BINSC_EPT; k++) slot[k] = v[k].x != 0xffffffffu ? atomicAdd(&cursor[v[k].x], 1u) : 0;
#pragma unroll
    for (int k = 0; k < BINSC_EPT; k++) if (v[k].x != 0xffffffffu) refs[slot[k]] = v[k].y;
}

// The same second level without a global atomic per reference (stands in for yrrid's in-shared-memory second pass,
// Partition4096.cu:352-433).  k_bin_scatter is bound by ~66 G global atomics-with-return per second (profiles/
// r01_ncu_scatter.txt); here a block of 1024 threads takes a tile of 64 K consecutive pairs -- all of one bin, bar the rare
// tile that straddles a boundary -- and a histogram of the bin's buckets in shared memory:
//   pass A  count the tile's pairs per bucket (shared-memory atomics, no return value needed);
//   reserve one global atomicAdd per bucket the tile touches (2 - 4 pairs each at 2^26 points) replaces the count by the first slot;
//   pass B  re-read the tile (L2), slot = shared-memory atomicAdd on that entry, store the reference.
// The lists a tile writes to are its bin's (<= 24 MB), and the ~150 tiles in flight span a handful of bins: L2-resident.
#define BS2_TPB 1024
#define BS2_TILE (64u * 1024u)     // measured at 2^26: 32 K 10.3 ms, 64 K 10.0, 128 K 10.5, 256 K 21.9 (the re-read falls out of L2)
#define BS2_MAX_SHIFT 15            // 32768 buckets per bin = 128 KB of shared memory
__global__ void __launch_bounds__(BS2_TPB) k_bin_scatter_tiled(const uint2 *__restrict__ staging, const uint32_t *__restrict__ bucket_off, uint32_t nbuckets,
                                                               int bin_shift, uint32_t tile, uint32_t *__restrict__ cursor, uint32_t *__restrict__ refs) {
    extern __shared__ uint32_t bs2_hist[];
    const uint32_t total = bucket_off[nbuckets];
    const uint64_t t0 = (uint64_t)blockIdx.x * tile;
    if (t0 >= total) return;
    const uint32_t t1 = total - t0 < tile ? total : (uint32_t)t0 + tile;
    const uint32_t per_bin = 1u << bin_shift, mask = per_bin - 1;
    uint32_t pos = (uint32_t)t0;
    while (pos < t1) {
        const uint32_t bin = staging[pos].x >> bin_shift;                       // the same word for every thread
        const uint32_t first = bin << bin_shift, next = first + per_bin < nbuckets ? first + per_bin : nbuckets;
        const uint32_t bin_end = bucket_off[next];                               // where this bin's pairs end in the staging array
        const uint32_t end = bin_end < t1 ? bin_end : t1;
        for (uint32_t b = threadIdx.x; b < per_bin; b += BS2_TPB) bs2_hist[b] = 0;
        __syncthreads();
        // (both passes: a warp whose lanes all hold the same key -- skewed scalar sets -- does one shared-memory atomic for the warp)
        for (uint32_t i = pos + threadIdx.x; i < end; i += BS2_TPB) {
            const uint32_t b = staging[i].x & mask, act = __activemask(), lead = __ffs(act) - 1;
            if (__all_sync(act, b == __shfl_sync(act, b, lead))) { if ((threadIdx.x & 31) == lead) atomicAdd(&bs2_hist[b], (uint32_t)__popc(act)); }
            else atomicAdd(&bs2_hist[b], 1u);
        }
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < per_bin; b += BS2_TPB) {
            const uint32_t c = bs2_hist[b];
            if (c) bs2_hist[b] = atomicAdd(&cursor[first + b], c);
        }
        __syncthreads();
        for (uint32_t i = pos + threadIdx.x; i < end; i += BS2_TPB) {
            const uint2 v = __ldcs(staging + i);
            const uint32_t b = v.x & mask, act = __activemask(), lead = __ffs(act) - 1;
            uint32_t slot;
            if (__all_sync(act, b == __shfl_sync(act, b, lead))) {
                uint32_t first = 0;
                if ((threadIdx.x & 31) == lead) first = atomicAdd(&bs2_hist[b], (uint32_t)__popc(act));
                slot = __shfl_sync(act, first, lead) + __popc(act & ((1u << (threadIdx.x & 31)) - 1));
            } else slot = atomicAdd(&bs2_hist[b], 1u);
            refs[slot] = v.y;
        }
        __syncthreads();
        pos = end;
    }
}

// ---------------------------------------------------------------------------------------------
// multi-GPU bucket reduce-scatter: the two cross-GPU counters (protocol at k_reduce_slices, kernels_cold.cuh)
// slot = this GPU's rank: flags[slot] of every peer (peer_flags[p] = peer p's flag array, mapped) <- epoch
__global__ void k_peer_signal(uint32_t *const *__restrict__ peer_flags, uint32_t world, uint32_t slot, uint32_t epoch) {
    if (threadIdx.x < world) {
        __threadfence_system();
        st_release_sys_c(peer_flags[threadIdx.x] + slot, epoch);
    }
}
// returns when every one of flags[0 .. world) has reached epoch (wrap-safe comparison)
__global__ void k_peer_wait(const uint32_t *__restrict__ flags, uint32_t world, uint32_t epoch) {
    if (threadIdx.x < world)
        while ((int32_t)(ld_acquire_sys_c(flags + threadIdx.x) - epoch) < 0) __nanosleep(100);
}

// ---------------------------------------------------------------------------------------------
// a11 (curve-independent part)
// replicate the first `have` entries of a table (vec 16-byte chunks each) up to n (vector doubling, as the reference
// generator does)
__global__ void __launch_bounds__(256) k_replicate(uint4 *table, uint32_t have, uint32_t n, uint32_t vec) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // one uint4 per thread
    size_t total = (size_t)(n - have) * vec;
    if (idx >= total) return;
    size_t e = have + idx / vec, k = idx % vec;
    table[e * vec + k] = table[(e % have) * vec + k];
}

}  // namespace b200msm

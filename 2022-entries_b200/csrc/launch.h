// launch.h -- host-callable launchers of the curve-templated kernels.  Each curve's hot and cold kernels live in
// their own translation unit (inst_*.cu) so the library builds in parallel; msm.cu only sees these declarations.
#pragma once
#include <cuda_runtime.h>
#include "grp.cuh"

namespace b200msm {

#define FIN_MAX_TERMS 24
struct fin_term_t { const xyzz_t *arr; uint32_t n; uint32_t stride_f; uint32_t shift; uint32_t kind; uint32_t base; };      // kind 1: element k weighs base + k
// one source of bucket sums for the level-0 reduction: a chunk of this GPU's points, or (multi-GPU bucket reduce-scatter) a
// chunk of a PEER GPU's points read over NVLink through a mapped pointer
struct red_src_t { const xyzz_t *res; const uint32_t *poff; };
#define RED_MAX_SRC 64
#define PEER_MAX 32
struct fin_args_t { fin_term_t term[FIN_MAX_TERMS]; int nterms; int F; int c; };
#define NORM_BATCH 16
#ifndef ACC_TPB
#define ACC_TPB 128
#endif
#ifndef ACC_MINB
#define ACC_MINB 3
#endif

template <class C> struct LaunchHot {
    static void digits(int pass, cudaStream_t st, const uint32_t *scalars, const uint32_t *infmask, uint32_t n, uint32_t first, uint32_t table_stride,
                       int c, int W, int F, uint32_t *counters, uint32_t *refs);
    static void digits_bin(cudaStream_t st, const uint32_t *scalars, const uint32_t *infmask, uint32_t n, uint32_t first, uint32_t table_stride,
                           int c, int W, int F, int bin_shift, uint32_t nbins, uint32_t *bin_cursor, void *staging);
    static void accumulate(cudaStream_t st, uint32_t max_pieces, const void *tables, const uint32_t *refs, const uint32_t *order,
                           const uint32_t *piece_start, const uint32_t *piece_len, const uint32_t *npieces_ptr, xyzz_t *res, int mix);
};
template <class C> struct LaunchCold {
    // n_out threads in blocks of `block` (n_out % block == 0); outT gets one entry per block
    // level 0 (srcs != nullptr): bucket k = bucket_base + output * m + j is the sum over nsrc sources of its pieces; with
    // wait_flags the kernel first waits until wait_n peers have published `epoch` (their bucket sums are complete)
    static void reduce_slices(cudaStream_t st, const xyzz_t *in, const red_src_t *srcs, uint32_t nsrc, uint32_t bucket_base, uint32_t n_out, uint32_t block, uint32_t m,
                              xyzz_t *outS, xyzz_t *outT, const uint32_t *wait_flags, uint32_t wait_n, uint32_t epoch);

    static void block_sum(cudaStream_t st, const xyzz_t *in, uint32_t ngroups, uint32_t group, xyzz_t *out);
    static void peer_merge(cudaStream_t st, const red_src_t *srcs, uint32_t nsrc, uint32_t bucket_base, uint32_t nslice, xyzz_t *merged,
                           const uint32_t *wait_flags, uint32_t wait_n, uint32_t epoch);
    static void finalize(cudaStream_t st, const fin_args_t &a, xyzz_t *out_xyzz, uint32_t *out_jac, xyzz_t *scratch, uint32_t *ticket);
    static void combine(cudaStream_t st, const xyzz_t *parts, int count, uint32_t *out_jac);
    static void double_n(cudaStream_t st, const affine_t *src, uint32_t n, int k, xyzz_t *tmp);
    static void batch_normalize(cudaStream_t st, const xyzz_t *tmp, uint32_t n, affine_t *dst, const uint32_t *infmask, uint32_t *bad);
    static void gen_points(cudaStream_t st, uint32_t first, uint32_t n, uint64_t seed, int mode, uint32_t distinct, const fp &gx, const fp &gy, xyzz_t *tmp);
    static void gen_scalars(cudaStream_t st, uint32_t n, uint64_t seed, uint32_t *out);
    static void convert_points(cudaStream_t st, const uint8_t *src, uint32_t n, uint32_t stride, affine_t *table0, uint32_t *infmask);
    static void export_points(cudaStream_t st, const void *entries, uint32_t n, uint32_t *dst, bool ark);
    static void make_entries(cudaStream_t st, const affine_t *src, uint32_t n, void *dst_entries, bool ark, const uint32_t *infmask, uint32_t *bad);
    static void check_order_r(cudaStream_t st, const void *entries, const uint32_t *infmask, uint32_t n, bool ark, uint32_t *outside);
    static size_t entry_bytes();
    static void test_grp(int op, const affine_abi_t *a, const affine_abi_t *b, uint32_t *out104, uint32_t n, int reps);
    static void generator(fp &gx, fp &gy);      // host: the curve generator in the device representation
    static void test_fp(int op, const fp_abi *a, const fp_abi *b, fp_abi *r, uint32_t n);
    static void test_ec(int op, xyzz_abi_t *acc, const void *other, uint32_t n, int reps);
};

static inline uint32_t cdiv(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

}  // namespace b200msm

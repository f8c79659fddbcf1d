// launch_hot.inl -- definitions of LaunchHot<C>; included by inst_hot_*.cu which instantiate one curve each.
#include "launch.h"
#include "kernels_hot.cuh"

namespace b200msm {

template <class C>
void LaunchHot<C>::digits(int pass, cudaStream_t st, const uint32_t *scalars, const uint32_t *infmask, uint32_t n, uint32_t first, uint32_t table_stride,
                          int c, int W, int F, uint32_t *counters, uint32_t *refs) {
    if (pass == 0) k_digits<C, 0><<<cdiv(n, 256), 256, 0, st>>>(scalars, infmask, n, first, table_stride, c, W, F, counters, refs);
    else k_digits<C, 1><<<cdiv(n, 256), 256, 0, st>>>(scalars, infmask, n, first, table_stride, c, W, F, counters, refs);
}
template <class C>
void LaunchHot<C>::digits_bin(cudaStream_t st, const uint32_t *scalars, const uint32_t *infmask, uint32_t n, uint32_t first, uint32_t table_stride,
                              int c, int W, int F, int bin_shift, uint32_t nbins, uint32_t *bin_cursor, void *staging) {
    k_digits_bin<C><<<cdiv(n, 256), 256, 0, st>>>(scalars, infmask, n, first, table_stride, c, W, F, bin_shift, nbins, bin_cursor, (uint2 *)staging);
}
template <class C>
void LaunchHot<C>::accumulate(cudaStream_t st, uint32_t max_pieces, const void *tables, const uint32_t *refs, const uint32_t *order,
                              const uint32_t *piece_start, const uint32_t *piece_len, const uint32_t *npieces_ptr, xyzz_t *res, int mix) {
    const typename Grp<C>::entry_t *t = (const typename Grp<C>::entry_t *)tables;
    if (mix == 0 && C::HAS_TE) k_accumulate<C, true><<<cdiv(max_pieces, ACC_TPB), ACC_TPB, 0, st>>>(t, refs, order, piece_start, piece_len, npieces_ptr, res);
    else k_accumulate<C, false><<<cdiv(max_pieces, ACC_TPB), ACC_TPB, 0, st>>>(t, refs, order, piece_start, piece_len, npieces_ptr, res);
}

}  // namespace b200msm

// launch_cold.inl -- definitions of LaunchCold<C>; included by inst_cold_*.cu which instantiate one curve each.
#include "launch.h"
#include "kernels_cold.cuh"

namespace b200msm {

template <class C>
void LaunchCold<C>::reduce_slices(cudaStream_t st, const xyzz_t *in, const red_src_t *srcs, uint32_t nsrc, uint32_t bucket_base, uint32_t n_out, uint32_t block, uint32_t m,
                                  xyzz_t *outS, xyzz_t *outT, const uint32_t *wait_flags, uint32_t wait_n, uint32_t epoch) {
    if (srcs) k_reduce_slices<C, true><<<n_out / block, block, 0, st>>>(nullptr, srcs, nsrc, bucket_base, m, outS, outT, wait_flags, wait_n, epoch);
    else k_reduce_slices<C, false><<<n_out / block, block, 0, st>>>(in, nullptr, 0, 0, m, outS, outT, nullptr, 0, 0);
}

template <class C>
void LaunchCold<C>::peer_merge(cudaStream_t st, const red_src_t *srcs, uint32_t nsrc, uint32_t bucket_base, uint32_t nslice, xyzz_t *merged,
                               const uint32_t *wait_flags, uint32_t wait_n, uint32_t epoch) {
    constexpr int smem = 2 * PEER_TPB * (int)sizeof(xyzz_t);
    static bool configured = false;
    if (!configured) { cudaFuncSetAttribute(k_peer_merge<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); configured = true; }
    k_peer_merge<C><<<cdiv(nslice, PEER_TPB), PEER_TPB, smem, st>>>(srcs, nsrc, bucket_base, nslice, merged, wait_flags, wait_n, epoch);
}
template <class C> void LaunchCold<C>::block_sum(cudaStream_t st, const xyzz_t *in, uint32_t ngroups, uint32_t group, xyzz_t *out) {
    k_block_sum<C><<<ngroups, group < 128 ? group : 128, 0, st>>>(in, group, out);
}
template <class C>
void LaunchCold<C>::finalize(cudaStream_t st, const fin_args_t &a, xyzz_t *out_xyzz, uint32_t *out_jac, xyzz_t *scratch, uint32_t *ticket) {
    k_finalize<C><<<a.F, 32 * FIN_WARPS, 0, st>>>(a, out_xyzz, out_jac, scratch, ticket);
}
template <class C> void LaunchCold<C>::combine(cudaStream_t st, const xyzz_t *parts, int count, uint32_t *out_jac) {
    k_combine<C><<<1, 32, 0, st>>>(parts, count, out_jac);
}
template <class C> void LaunchCold<C>::double_n(cudaStream_t st, const affine_t *src, uint32_t n, int k, xyzz_t *tmp) {
    k_double_n<C><<<cdiv(n, 128), 128, 0, st>>>(src, n, k, tmp);
}
template <class C> void LaunchCold<C>::batch_normalize(cudaStream_t st, const xyzz_t *tmp, uint32_t n, affine_t *dst, const uint32_t *infmask, uint32_t *bad) {
    k_batch_normalize<C><<<cdiv(cdiv(n, NORM_BATCH), 128), 128, 0, st>>>(tmp, n, dst, infmask, bad);
}
template <class C>
void LaunchCold<C>::gen_points(cudaStream_t st, uint32_t first, uint32_t n, uint64_t seed, int mode, uint32_t distinct, const fp &gx, const fp &gy, xyzz_t *tmp) {
    k_gen_points<C><<<cdiv(n, 128), 128, 0, st>>>(first, n, seed, mode, distinct, gx, gy, tmp);
}
template <class C> void LaunchCold<C>::gen_scalars(cudaStream_t st, uint32_t n, uint64_t seed, uint32_t *out) {
    k_gen_scalars<C><<<cdiv(n, 256), 256, 0, st>>>(n, seed, out);
}
template <class C> void LaunchCold<C>::convert_points(cudaStream_t st, const uint8_t *src, uint32_t n, uint32_t stride, affine_t *table0, uint32_t *infmask) {
    k_convert_points<C><<<cdiv(n, 128), 128, 0, st>>>(src, n, stride, table0, infmask);
}
template <class C> void LaunchCold<C>::export_points(cudaStream_t st, const void *entries, uint32_t n, uint32_t *dst, bool ark) {
    k_export_points<C><<<cdiv(n, 128), 128, 0, st>>>((const typename Grp<C>::entry_t *)entries, n, dst, ark);
}
template <class C> void LaunchCold<C>::make_entries(cudaStream_t st, const affine_t *src, uint32_t n, void *dst_entries, bool ark, const uint32_t *infmask, uint32_t *bad) {
    k_make_entries<C><<<cdiv(cdiv(n, GRP_BATCH), 128), 128, 0, st>>>(src, n, (typename Grp<C>::entry_t *)dst_entries, ark, infmask, bad);
}
template <class C> void LaunchCold<C>::check_order_r(cudaStream_t st, const void *entries, const uint32_t *infmask, uint32_t n, bool ark, uint32_t *outside) {
    k_check_order_r<C><<<cdiv(n, 128), 128, 0, st>>>((const typename Grp<C>::entry_t *)entries, infmask, n, ark, outside);
}
template <class C> size_t LaunchCold<C>::entry_bytes() { return sizeof(typename Grp<C>::entry_t); }
template <class C> void LaunchCold<C>::test_grp(int op, const affine_abi_t *a, const affine_abi_t *b, uint32_t *out104, uint32_t n, int reps) {
    k_test_grp<C><<<cdiv(n, 64), 64>>>(op, a, b, out104, n, reps);
}
template <class C> void LaunchCold<C>::test_fp(int op, const fp_abi *a, const fp_abi *b, fp_abi *r, uint32_t n) {
    k_test_fp<C><<<cdiv(n, 128), 128>>>(op, a, b, r, n);
}
template <class C> void LaunchCold<C>::test_ec(int op, xyzz_abi_t *acc, const void *other, uint32_t n, int reps) {
    k_test_ec<C><<<cdiv(n, 64), 64>>>(op, acc, other, n, reps);
}

}  // namespace b200msm

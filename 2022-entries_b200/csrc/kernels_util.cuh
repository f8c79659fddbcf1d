// kernels_util.cuh -- device helpers shared by the kernel headers (async copies, 16-byte vector moves, RNG).
#pragma once
#include <cuda_runtime.h>
#include "grp.cuh"

namespace b200msm {

// ---------------------------------------------------------------------------------------------
// small helpers
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int AFF_VEC = sizeof(affine_t) / 16;     // 7 x 16 bytes
constexpr int XYZZ_VEC = sizeof(xyzz_t) / 16;      // 14 x 16 bytes
static_assert(sizeof(affine_t) == 112 && sizeof(xyzz_t) == 224, "device record sizes");

__device__ __forceinline__ void load_xyzz(xyzz_t &d, const xyzz_t *src) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *t = reinterpret_cast<uint4 *>(&d);
#pragma unroll
    for (int i = 0; i < XYZZ_VEC; i++) t[i] = s[i];
}
__device__ __forceinline__ void store_xyzz(xyzz_t *dst, const xyzz_t &v) {
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    const uint4 *t = reinterpret_cast<const uint4 *>(&v);
#pragma unroll
    for (int i = 0; i < XYZZ_VEC; i++) d[i] = t[i];
}
__device__ __forceinline__ void load_affine(affine_t &d, const affine_t *src) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *t = reinterpret_cast<uint4 *>(&d);
#pragma unroll
    for (int i = 0; i < AFF_VEC; i++) t[i] = __ldg(s + i);
}
__device__ __forceinline__ void store_affine(affine_t *dst, const affine_t &v) {
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    const uint4 *t = reinterpret_cast<const uint4 *>(&v);
#pragma unroll
    for (int i = 0; i < AFF_VEC; i++) d[i] = t[i];
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

}  // namespace b200msm

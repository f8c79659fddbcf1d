// Field-multiplier throughput: integer (fp.cuh, 14 x 28-bit IMAD.WIDE) vs FP64 (fpd.cuh, 8 x 48-bit DFMA), BLS12-377.
// Chains of dependent multiplies per thread (the ILP inside one multiply is what the accumulate kernel has too).
// Build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -o mul_bench mul_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../2022-entries_b200/csrc/fpd.cuh"
using namespace b200msm;
#define ITERS 512
template <int MODE, int TPB> __global__ void __launch_bounds__(TPB) k(uint32_t *io) {
    fp_abi a, b;
    for (int i = 0; i < 12; i++) { a.w[i] = io[(threadIdx.x & 127) * 12 + i]; b.w[i] = io[4096 + (threadIdx.x & 127) * 12 + i]; }
    a.w[11] &= 0xffffff; b.w[11] &= 0xffffff;
    if (MODE == 0) {
        fp x, y; Fp<Bls12_377>::repack_from_abi(x, a); Fp<Bls12_377>::repack_from_abi(y, b);
        for (int it = 0; it < ITERS; it++) { Fp<Bls12_377>::mul_schoolbook(x, x, y); Fp<Bls12_377>::mul_schoolbook(y, y, x); }
        Fp<Bls12_377>::repack_to_abi(a, x); Fp<Bls12_377>::repack_to_abi(b, y);
    } else if (MODE == 6) {
        fp x, y; Fp<Bls12_377>::repack_from_abi(x, a); Fp<Bls12_377>::repack_from_abi(y, b);
        for (int it = 0; it < ITERS; it++) { Fp<Bls12_377>::mul_karatsuba(x, x, y); Fp<Bls12_377>::mul_karatsuba(y, y, x); }
        Fp<Bls12_377>::repack_to_abi(a, x); Fp<Bls12_377>::repack_to_abi(b, y);
    } else if (MODE == 1) {
        fpd x, y; FpD<Bls12_377>::from_abi(x, a); FpD<Bls12_377>::from_abi(y, b);
        for (int it = 0; it < ITERS; it++) { FpD<Bls12_377>::mul(x, x, y); FpD<Bls12_377>::mul(y, y, x); }
        FpD<Bls12_377>::to_abi(a, x); FpD<Bls12_377>::to_abi(b, y);
    } else if (MODE >= 3) {   // warp-specialised mix: some warps on the integer pipe, the others on the FP64 pipe
        const int warp = threadIdx.x >> 5;
        const bool integer = MODE == 3 ? (warp & 1) : (MODE == 4 ? (warp % 3 == 0) : (warp % 4 == 0));
        if (integer) {
            fp x, y; Fp<Bls12_377>::repack_from_abi(x, a); Fp<Bls12_377>::repack_from_abi(y, b);
            for (int it = 0; it < ITERS; it++) { Fp<Bls12_377>::mul(x, x, y); Fp<Bls12_377>::mul(y, y, x); }
            Fp<Bls12_377>::repack_to_abi(a, x); Fp<Bls12_377>::repack_to_abi(b, y);
        } else {
            fpd x, y; FpD<Bls12_377>::from_abi(x, a); FpD<Bls12_377>::from_abi(y, b);
            for (int it = 0; it < ITERS; it++) { FpD<Bls12_377>::mul(x, x, y); FpD<Bls12_377>::mul(y, y, x); }
            FpD<Bls12_377>::to_abi(a, x); FpD<Bls12_377>::to_abi(b, y);
        }
    } else {        // two independent chains per thread (more ILP)
        fpd x, y, u, v; FpD<Bls12_377>::from_abi(x, a); FpD<Bls12_377>::from_abi(y, b); u = y; v = x;
        for (int it = 0; it < ITERS / 2; it++) { FpD<Bls12_377>::mul(x, x, y); FpD<Bls12_377>::mul(u, u, v); FpD<Bls12_377>::mul(y, y, x); FpD<Bls12_377>::mul(v, v, u); }
        FpD<Bls12_377>::add(x, x, u); FpD<Bls12_377>::add(y, y, v); FpD<Bls12_377>::norm(x); FpD<Bls12_377>::norm(y);
        FpD<Bls12_377>::to_abi(a, x); FpD<Bls12_377>::to_abi(b, y);
    }
    uint32_t s = 0;
    for (int i = 0; i < 12; i++) s ^= a.w[i] ^ b.w[i];
    if (s == 0x12345678) io[0] = s;
}
template <int MODE, int TPB> void run(const char *name, int blocks_per_sm, uint32_t *io) {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = sms * blocks_per_sm;
    k<MODE, TPB><<<grid, TPB>>>(io); cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); k<MODE, TPB><<<grid, TPB>>>(io); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double muls = (double)grid * TPB * ITERS * 2;
    printf("%-34s warps/SM=%2d  %.3f ms  %.2f G mul/s  %.2f SM-clocks per thread-multiply (at 1965 MHz)  err=%s\n", name, blocks_per_sm * TPB / 32, best,
           muls / (best * 1e-3) / 1e9, 1.965e9 * sms / (muls / (best * 1e-3)), cudaGetErrorString(cudaGetLastError()));
}
int main() {
    uint32_t *io; cudaMalloc(&io, 65536); cudaMemset(io, 0x5a, 65536);
    for (int b : {1, 2, 4}) {
        run<0, 128>("integer 14x28 (IMAD.WIDE)", b, io);
        run<6, 128>("integer 14x28 Karatsuba (329 IMAD)", b, io);
        run<1, 128>("fp64 8x48 (DFMA)", b, io);
        run<4, 384>("mix 1 int : 2 fp64 warps (x3 TPB)", b, io);
    }
    return 0;
}

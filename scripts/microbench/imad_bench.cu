// SUPERSEDED by mul_bench.cu: the loop-invariant products below were partly strength-reduced by ptxas, so the
// IMAD.WIDE figure this printed (63 / clk / SM) is twice the real rate (32; profiles/r01_mul_bench.txt).  Kept because
// profiles/r01_imad_microbench.txt cites it.
// Integer-pipe microbenchmark (measure, don't assume): issue rate of the instructions the Montgomery multiply is
// made of, per SM per clock.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_bench imad_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE> __global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t a0, uint32_t b0) {
    uint32_t a = a0 + threadIdx.x, b = b0 + blockIdx.x;
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { lo[i] = i + a; hi[i] = i * 3 + b; }
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {          // 8 independent IMAD.WIDE.U32 (64-bit accumulate)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint64_t acc = ((uint64_t)hi[i] << 32) | lo[i];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
                lo[i] = (uint32_t)acc; hi[i] = (uint32_t)(acc >> 32);
            }
        } else if (MODE == 1) {   // 16 independent 32-bit IMAD (mad.lo)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(a), "r"(b));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi[i]) : "r"(b), "r"(a));
            }
        } else if (MODE == 2) {   // two carry chains of 4 IMAD.WIDE.U32.X each (the shape inside the multiplier)
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %8, %9, %2; madc.hi.cc.u32 %3, %8, %9, %3;"
                         "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5; madc.lo.cc.u32 %6, %8, %9, %6; madc.hi.u32 %7, %8, %9, %7;"
                         : "+r"(lo[0]), "+r"(lo[1]), "+r"(lo[2]), "+r"(lo[3]), "+r"(lo[4]), "+r"(lo[5]), "+r"(lo[6]), "+r"(lo[7]) : "r"(a), "r"(b));
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %8, %9, %2; madc.hi.cc.u32 %3, %8, %9, %3;"
                         "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5; madc.lo.cc.u32 %6, %8, %9, %6; madc.hi.u32 %7, %8, %9, %7;"
                         : "+r"(hi[0]), "+r"(hi[1]), "+r"(hi[2]), "+r"(hi[3]), "+r"(hi[4]), "+r"(hi[5]), "+r"(hi[6]), "+r"(hi[7]) : "r"(b), "r"(a));
        } else if (MODE == 3) {   // 8 IMAD.WIDE + 8 independent IADD3-class adds (does the ALU pipe issue alongside?)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint64_t acc = ((uint64_t)hi[i] << 32) | lo[i];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
                lo[i] = (uint32_t)acc; hi[i] = (uint32_t)(acc >> 32);
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(lo[(i + 3) & 7]));
            }
        } else if (MODE == 4) {   // 8 DFMA-free check: 8 mul.hi.u32 + 8 mul.lo (separate lo/hi halves: 2 IMAD per product)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(hi[i]) : "r"(a), "r"(lo[i]));
            }
        } else if (MODE == 5) {   // FP64 DFMA rate (for a possible double-precision multiplier path)
            double *d = reinterpret_cast<double *>(lo);   // 4 doubles in lo[], 4 in hi[]
            double *e = reinterpret_cast<double *>(hi);
            double x = __longlong_as_double(0x3ff0000000000001ULL + a), y = 1e-9;
#pragma unroll
            for (int i = 0; i < 4; i++) { asm volatile("fma.rn.f64 %0, %1, %0, %2;" : "+d"(d[i]) : "d"(x), "d"(y)); asm volatile("fma.rn.f64 %0, %1, %0, %2;" : "+d"(e[i]) : "d"(x), "d"(y)); }
        }
    }
    uint32_t s = a;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= lo[i] ^ hi[i];
    if (s == 0x12345678) out[0] = s;
}
template <int MODE> void run(const char *name, double ops_per_iter, int blocks_per_sm) {
    int dev = 0, sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    uint32_t *out; cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = sms * blocks_per_sm;
    k<MODE><<<grid, 256>>>(out, 3, 5); cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); k<MODE><<<grid, 256>>>(out, 3, 5); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double total = (double)grid * 256 * ITERS * ops_per_iter;
    double per_s = total / (best * 1e-3);
    printf("%-44s blocks/SM=%d  %.3f ms  %.2f Tops/s  = %.1f ops/clk/SM at %d MHz nominal (%d SMs)\n", name, blocks_per_sm, best, per_s / 1e12, per_s / sms / (clk * 1e3), clk / 1000, sms);
    cudaFree(out);
}
int main() {
    for (int b : {2, 4, 8}) {
        run<0>("IMAD.WIDE.U32 independent", 8, b);
        run<1>("IMAD (32-bit) independent", 16, b);
        run<2>("IMAD.WIDE.U32.X carry chains (2 x 4)", 8, b);
        run<3>("IMAD.WIDE.U32 + IADD (8 + 8)", 8, b);
        run<4>("IMAD.HI.U32", 8, b);
        run<5>("DFMA", 8, b);
    }
    return 0;
}

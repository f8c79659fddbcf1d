// Pipe-rate microbenchmark v2 (supersedes imad_bench.cu, whose loop-invariant products ptxas partly strength-reduced).
// Every operand depends on loop-carried state, so nothing can be hoisted; check the SASS loop body with
//   cuobjdump -sass pipe_bench | grep -c <mnemonic>
// Reports warp-instructions per clock per SM for each instruction class and for the mixes the field multiplier
// could be built from.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench pipe_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define NCH 8
template <int MODE> __global__ void __launch_bounds__(256) k(uint64_t *out, const uint32_t *in, uint32_t b0) {
    uint64_t acc[NCH], sum[NCH];
    double d[NCH], da[NCH];
    uint32_t a[NCH], w[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) {
        a[i] = in[threadIdx.x + 32 * i]; w[i] = a[i] ^ b0; acc[i] = a[i]; sum[i] = i;
        da[i] = 1.0 + 1e-9 * a[i]; d[i] = 0.5 * i;
    }
    for (int it = 0; it < ITERS; it++) {
        // one multiplier per iteration, taken from last iteration's results: the shape of a Montgomery round
        const uint32_t m = (uint32_t)acc[NCH - 1] ^ w[NCH - 1] ^ (uint32_t)sum[NCH - 1];
        const double dm = d[NCH - 1];
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            if (MODE == 0 || MODE == 4 || MODE == 5 || MODE == 6)         // IMAD.WIDE.U32: acc_i += a_i * m
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a[i]), "r"(m));
            if (MODE == 1)                                                 // 32-bit IMAD: w_i += a_i * m
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(w[i]) : "r"(a[i]), "r"(m));
            if (MODE == 2 || MODE == 4 || MODE == 6 || MODE == 7)         // DFMA: d_i += da_i * dm
                asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(da[i]), "d"(dm));
            if (MODE == 3 || MODE == 5 || MODE == 6 || MODE == 7) {       // 64-bit add (IADD3 + IADD3.X): sum_i += x
                uint64_t x = (MODE == 7 || MODE == 6) ? (uint64_t)__double_as_longlong(d[i]) : (acc[i] ^ (uint64_t)m);
                asm volatile("{ .reg .u32 l0, h0, l1, h1; mov.b64 {l0, h0}, %0; mov.b64 {l1, h1}, %1; add.cc.u32 l0, l0, l1; addc.u32 h0, h0, h1; mov.b64 %0, {l0, h0}; }" : "+l"(sum[i]) : "l"(x));
            }
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s ^= acc[i] ^ sum[i] ^ (uint64_t)__double_as_longlong(d[i]) ^ w[i];
    if (s == 0x12345678) out[0] = s;
}
template <int MODE> void run(const char *name, int blocks_per_sm) {
    int dev = 0, sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t *out; cudaMalloc(&out, 8); uint32_t *in; cudaMalloc(&in, 4096); cudaMemset(in, 0x5a, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = sms * blocks_per_sm;
    k<MODE><<<grid, 256>>>(out, in, 5); cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); k<MODE><<<grid, 256>>>(out, in, 5); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    // loop iterations of the whole grid per second; one iteration = NCH "slots" per thread
    double slots = (double)grid * 256 * ITERS * NCH;
    double clk = 1.965e9;
    printf("%-58s blocks/SM=%d  %.3f ms  %.2f thread-slots/clk/SM (at 1965 MHz)\n", name, blocks_per_sm, best, slots / (best * 1e-3) / sms / clk);
    cudaFree(out);
}
int main() {
    for (int b : {2, 4}) {
        run<0>("IMAD.WIDE.U32 (1 per slot)", b);
        run<1>("IMAD 32-bit lo (1 per slot)", b);
        run<2>("DFMA (1 per slot)", b);
        run<3>("64-bit add = IADD3 + IADD3.X (2 ALU per slot)", b);
        run<4>("IMAD.WIDE + DFMA (1 + 1 per slot)", b);
        run<5>("IMAD.WIDE + 64-bit add (1 + 2 per slot)", b);
        run<6>("IMAD.WIDE + DFMA + 64-bit add (1 + 1 + 2 per slot)", b);
        run<7>("DFMA + 64-bit add of its bits (1 + 2 per slot)", b);
    }
    return 0;
}

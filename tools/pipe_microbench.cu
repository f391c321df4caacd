// Issue-rate probe for the integer / fp64 pipes of sm_100a (run under gpurun):
// per-SM throughput (thread-ops per clock) of mad.lo, mad.hi, mad.wide, their carry-chain forms, DFMA,
// IADD3, and mixes — decides how the Montgomery multiplier should be written.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ACC 8

template <int MODE> __global__ void k(uint32_t* out, uint32_t a0, uint32_t b0, double d0) {
    uint32_t a = a0 + threadIdx.x, b = b0 + blockIdx.x;
    uint32_t x[ACC]; unsigned long long w[ACC]; double f[ACC];
#pragma unroll
    for (int i = 0; i < ACC; i++) { x[i] = i + a; w[i] = i + b; f[i] = d0 + i; }
    double fa = d0 * 1.0000001, fb = d0 * 0.9999999;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ACC; i++) {
            if (MODE == 0) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a), "r"(b));
            if (MODE == 1) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a), "r"(b));
            if (MODE == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a), "r"(b));
            if (MODE == 3) asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(f[i]) : "d"(fa), "d"(fb));
            if (MODE == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
            if (MODE == 5) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a), "r"(b));
                             asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(f[i]) : "d"(fa), "d"(fb)); }
            if (MODE == 6) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a), "r"(b));
                             asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a)); }
            if (MODE == 7) { asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a), "r"(b));
                             asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x[(i + 1) % ACC]) : "r"(a), "r"(b)); }
            if (MODE == 8) asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) % ACC]));
            if (MODE == 9) { asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a), "r"(b));     // carry chain pair (fuses to WIDE?)
                             asm volatile("madc.hi.u32 %0, %1, %2, %0;" : "+r"(x[(i + 1) % ACC]) : "r"(a), "r"(b)); }
            if (MODE == 10) { asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(f[i]) : "d"(fa), "d"(fb));
                              asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) % ACC])); }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ACC; i++) s += x[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32) + (uint32_t)f[i];
    if (s == 0x12345u) out[threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int ops_per_slot) {
    uint32_t* out; cudaMalloc(&out, 4096);
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wpsm : {16, 32, 64}) {
        int tpb = 256, blocks = sm * wpsm * 32 / tpb;
        k<MODE><<<blocks, tpb>>>(out, 3, 5, 1.5); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<MODE><<<blocks, tpb>>>(out, 3, 5, 1.5); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)blocks * tpb * ITERS * ACC * ops_per_slot;
        printf("%-28s warps/SM=%2d  %.3f ms  %.1f thread-ops/clk/SM (at %.0f MHz nominal)  %.3e ops/s\n", name, wpsm, ms,
               ops / (ms * 1e-3) / sm / (clk * 1e3), clk / 1e3, ops / (ms * 1e-3));
    }
}
int main() {
    run<0>("mad.lo.u32", 1); run<1>("mad.hi.u32", 1); run<2>("mad.wide.u32", 1); run<3>("fma.rz.f64", 1);
    run<4>("add.u32", 1); run<8>("add.u64", 1); run<7>("mad.lo + mad.hi (2 ops)", 2); run<9>("mad.lo.cc+madc.hi (2 ops)", 2);
    run<5>("mad.wide + dfma (2 ops)", 2); run<6>("mad.wide + add.u32 (2 ops)", 2); run<10>("dfma + add.u64 (2 ops)", 2);
    return 0;
}

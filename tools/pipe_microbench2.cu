// Second issue-rate probe: per-thread (non-uniform) operands, SASS-checked instruction forms.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define ACC 8
template <int MODE> __global__ void k(uint32_t* out, const uint32_t* in) {
    uint32_t a[ACC], b[ACC]; unsigned long long w[ACC]; uint32_t x[ACC], y[ACC];
#pragma unroll
    for (int i = 0; i < ACC; i++) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 32 * i + 7]; w[i] = a[i]; x[i] = b[i]; y[i] = a[i] ^ b[i]; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ACC; i++) {
            if (MODE == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[(i + 3) % ACC]), "r"(b[i]));           // WIDE RRR
            if (MODE == 1) asm volatile("mad.wide.u32 %0, %1, 0x0800170b, %0;" : "+l"(w[i]) : "r"((uint32_t)w[(i + 3) % ACC]));                            // WIDE R,imm
            if (MODE == 2) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(x[(i + 3) % ACC]), "r"(b[i]));              // IMAD lo RRR
            if (MODE == 3) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(x[(i + 3) % ACC]), "r"(b[i]));              // IMAD.HI RRR
            if (MODE == 4) { asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(x[i]), "+r"(y[i]) : "r"(x[(i + 3) % ACC]), "r"(b[i])); }   // fused pair?
            if (MODE == 5) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((uint32_t)(w[(i + 3) % ACC] >> 32)), "r"(b[i])); // WIDE with RZ addend (dep via xor)
            if (MODE == 6) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(a[i]), "r"(b[i]));   // 64-bit add as IADD3 + IADD3.X
            if (MODE == 7) asm volatile("mad.lo.u32 %0, %1, 0x0800170b, %0;" : "+r"(x[i]) : "r"(x[(i + 3) % ACC]));                              // IMAD lo R,imm
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ACC; i++) s += x[i] + y[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
    if (s == 0x12345u) out[threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int ops) {
    uint32_t *out, *in; cudaMalloc(&out, 4096); cudaMalloc(&in, 8192); cudaMemset(in, 0x3c, 8192);
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wpsm : {8, 32}) {
        int tpb = 256, blocks = sm * wpsm * 32 / tpb;
        k<MODE><<<blocks, tpb>>>(out, in); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<MODE><<<blocks, tpb>>>(out, in); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double n = (double)blocks * tpb * ITERS * ACC * ops;
        printf("%-34s warps/SM=%2d  %.1f thread-ops/clk/SM (1965 MHz)  => %.2f clk per warp-op per SMSP\n", name, wpsm,
               n / (ms * 1e-3) / sm / 1.965e9, 128.0 / (n / (ms * 1e-3) / sm / 1.965e9));
    }
}
int main() {
    run<0>("IMAD.WIDE.U32 R,R,R64", 1); run<1>("IMAD.WIDE.U32 R,imm,R64", 1); run<2>("IMAD R,R,R (lo)", 1); run<7>("IMAD R,imm,R (lo)", 1);
    run<3>("IMAD.HI.U32 R,R,R", 1); run<4>("mad.lo.cc+madc.hi pair", 1); run<5>("IMAD.WIDE.U32 R,R,RZ", 1); run<6>("IADD3+IADD3.X (64-bit add)", 1);
    return 0;
}

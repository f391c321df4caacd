// Can B200 run an FP64-FMA big-integer multiplier NEXT TO the IMAD one?  Kernel A runs the production Fq
// multiplication (IMAD.WIDE carry chains, fmaheavy pipe).  Kernel B runs the instruction mix of an Emmart-style
// 52-bit-limb product phase: per limb product 2 DFMA + 1 DADD (FP64 pipe) + 2 int64 accumulations (ALU pipe),
// 128 limb products per "multiplication" (8x8 for a·b + 8x8 for m·p).  A alone, B alone, then A and B on two
// streams sharing every SM; reports the rates.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dual experiments/dual_pipe/dual_pipe_microbench.cu
#include <cstdio>
#include <cstdint>
#include "../../snarkvm_b200/csrc/ff.cuh"
using namespace b200;

__global__ void __launch_bounds__(128) k_imad(const uint32_t* in, uint32_t* out, int iters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    Fq x = Fq::load(in + 12 * (i % 256)), y = Fq::load(in + 12 * ((i + 7) % 256));
    for (int it = 0; it < iters; it++) { x = x * y; y = y * x; }
    if (x.v[0] + y.v[0] == 0x1234567u) x.store(out);
}
__global__ void __launch_bounds__(128) k_dfma(const uint32_t* in, uint32_t* out, int iters) {
    const double C1 = 20282409603651670423947251286016.0;            // 2^104
    const double C2 = 20282409603651674927546878656512.0;            // 2^104 + 2^52
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double a[8], b[8], m[8], p[8]; long long acc[17];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        a[k] = (double)(in[(i + k) % 256] & 0xfffff) * 4294967296.0 + in[(i + 9 * k) % 256];
        b[k] = (double)(in[(i + 3 * k) % 256] & 0xfffff) * 4294967296.0 + 7;
        m[k] = (double)(in[(i + 5 * k) % 256] & 0xfffff) * 4294967296.0 + 11;
        p[k] = (double)(in[(i + 7 * k) % 256] & 0xfffff) * 4294967296.0 + 13;
    }
#pragma unroll
    for (int k = 0; k < 17; k++) acc[k] = k;
    for (int it = 0; it < 2 * iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                double h = __fma_rz(a[r], b[j], C1), s = C2 - h, l = __fma_rz(a[r], b[j], s);
                acc[r + j] += __double_as_longlong(l); acc[r + j + 1] += __double_as_longlong(h);
            }
            // the reduction row depends on the running column, like the real thing
            m[r] = (double)(acc[r] & 0xfffffffffffffLL);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                double h = __fma_rz(m[r], p[j], C1), s = C2 - h, l = __fma_rz(m[r], p[j], s);
                acc[r + j] += __double_as_longlong(l); acc[r + j + 1] += __double_as_longlong(h);
            }
        }
        a[it & 7] = (double)(acc[8 + (it & 7)] & 0xfffffffffffffLL);
    }
    long long s = 0;
#pragma unroll
    for (int k = 0; k < 17; k++) s += acc[k];
    if (s == 0x1234567) out[1] = (uint32_t)s;
}
int main() {
    uint32_t *in, *out; cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 4096); cudaMemset(in, 0x5a, 4096 * 4);
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    const int iters = 1000;
    for (int wa : {8, 16}) for (int wb : {8, 16}) {        // warps per SM given to each kernel (128-thread CTAs)
        int ba = sm * wa / 4, bb = sm * wb / 4; float ta, tb, tab;
        k_imad<<<ba, 128>>>(in, out, 4); k_dfma<<<bb, 128>>>(in, out, 4); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_imad<<<ba, 128>>>(in, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ta, e0, e1);
        cudaEventRecord(e0); k_dfma<<<bb, 128>>>(in, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&tb, e0, e1);
        cudaDeviceSynchronize();
        cudaEventRecord(e0, s1);
        cudaStreamWaitEvent(s2, e0, 0);
        k_imad<<<ba, 128, 0, s1>>>(in, out, iters);
        k_dfma<<<bb, 128, 0, s2>>>(in, out, iters);
        cudaEventRecord(e2, s2); cudaStreamWaitEvent(s1, e2, 0); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&tab, e0, e1);
        double na = (double)ba * 128 * iters * 2, nb = (double)bb * 128 * iters * 2;
        printf("IMAD warps/SM=%2d DFMA warps/SM=%2d: alone %.3f ms (%.3e mul/s) | alone %.3f ms (%.3e mul-like/s) | together %.3f ms -> %.3e + %.3e = %.3e (x%.2f of IMAD alone)\n",
               wa, wb, ta, na / (ta * 1e-3), tb, nb / (tb * 1e-3), tab, na / (tab * 1e-3), nb / (tab * 1e-3), (na + nb) / (tab * 1e-3),
               (na + nb) / (tab * 1e-3) / (na / (ta * 1e-3)));
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

// Throughput of the radix-2^28 Fq multiplier (inlined vs called) next to the saturated 12x32 one.
#include <cstdio>
#include "fq28.cuh"
using namespace b200;
__global__ void k_chain28(const uint32_t* a, uint32_t* c, int iters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    Fq28 x, y;
    for (int k = 0; k < 14; k++) { x.v[k] = a[(14 * i + k) % 4096] & 0x0fffffff; y.v[k] = a[(14 * i + k + 14) % 4096] & 0x0fffffff; }
    x.v[13] &= 0xfff; y.v[13] &= 0xfff;
    for (int it = 0; it < iters; it++) { x = fq28_mul(x, y); y = fq28_mul(y, x); }
    uint32_t s = 0; for (int k = 0; k < 14; k++) s += x.v[k] + y.v[k];
    if (s == 0x1234567u) c[i % 4096] = s;
}
__global__ void k_chain28_sqr(const uint32_t* a, uint32_t* c, int iters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    Fq28 x, y;
    for (int k = 0; k < 14; k++) { x.v[k] = a[(14 * i + k) % 4096] & 0x0fffffff; y.v[k] = a[(14 * i + k + 14) % 4096] & 0x0fffffff; }
    x.v[13] &= 0xfff; y.v[13] &= 0xfff;
    for (int it = 0; it < iters; it++) { x = fq28_sqr(x); y = fq28_sqr(y); }
    uint32_t s = 0; for (int k = 0; k < 14; k++) s += x.v[k] + y.v[k];
    if (s == 0x1234567u) c[i % 4096] = s;
}
__global__ void k_chain32(const uint32_t* a, uint32_t* c, int iters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    Fq x = Fq::load(a + 12 * (i % 256)), y = Fq::load(a + 12 * ((i + 7) % 256));
    for (int it = 0; it < iters; it++) { x = x * y; y = y * x; }
    if (x.v[0] + y.v[0] == 0x1234567u) x.store(c);
}
int main() {
    uint32_t *a, *c; cudaMalloc(&a, 4096 * 4 + 64); cudaMalloc(&c, 4096 * 4 + 64);
    cudaMemset(a, 0x5a, 4096 * 4 + 64);
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1000;
    for (int tpb : {128, 256}) for (int bps : {1, 2, 4, 8}) {
        int blocks = sm * bps; float ms;
        k_chain28<<<blocks, tpb>>>(a, c, 4); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_chain28<<<blocks, tpb>>>(a, c, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        double r28 = 2.0 * blocks * tpb * iters / (ms * 1e-3);
        cudaEventRecord(e0); k_chain28_sqr<<<blocks, tpb>>>(a, c, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        double s28 = 2.0 * blocks * tpb * iters / (ms * 1e-3);
        k_chain32<<<blocks, tpb>>>(a, c, 4); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_chain32<<<blocks, tpb>>>(a, c, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        double r32 = 2.0 * blocks * tpb * iters / (ms * 1e-3);
        printf("tpb=%d blocks/SM=%d warps/SM=%d : fq28 mul %.3e/s  fq28 sqr %.3e/s  fq32 mul %.3e/s  (%s)\n", tpb, bps, tpb * bps / 32, r28, s28, r32,
#ifdef FQ28_INLINE_MUL
               "inlined"
#else
               "called"
#endif
        );
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

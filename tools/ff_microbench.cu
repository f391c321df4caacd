// Standalone sanity + throughput probe for the device field arithmetic (run under gpurun).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gpurun_out/ff_microbench tools/ff_microbench.cu
// 1. checks Fr/Fq mul/add/sub/inverse on random inputs against a host CIOS reference written here
// 2. times dependent-chain and independent-stream multiplications to get Fr/Fq mults per second.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../snarkvm_b200/csrc/ff.cuh"
using namespace b200;
typedef unsigned __int128 u128;

template <int N64> static void host_mont_mul(uint64_t* r, const uint64_t* a, const uint64_t* b, const uint64_t* m, uint64_t inv) {
    uint64_t t[N64 + 2]; memset(t, 0, sizeof t);
    for (int i = 0; i < N64; i++) {
        u128 c = 0;
        for (int j = 0; j < N64; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[N64]; t[N64] = (uint64_t)c; t[N64 + 1] = (uint64_t)(c >> 64);
        uint64_t k = t[0] * inv; c = (u128)k * m[0] + t[0]; c >>= 64;
        for (int j = 1; j < N64; j++) { c += (u128)k * m[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[N64]; t[N64 - 1] = (uint64_t)c; t[N64] = t[N64 + 1] + (uint64_t)(c >> 64);
    }
    bool ge = t[N64] != 0;
    if (!ge) { ge = true; for (int i = N64 - 1; i >= 0; i--) { if (t[i] != m[i]) { ge = t[i] > m[i]; break; } } }
    if (ge) { uint64_t br = 0; for (int i = 0; i < N64; i++) { u128 d = (u128)t[i] - m[i] - br; t[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; } }
    memcpy(r, t, 8 * N64);
}

template <class F> __global__ void k_sqr_check(const uint32_t* a, uint32_t* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    F x = F::load(a + F::N * i);
    F d = x.sqr() - x * x;
    d.store(out + F::N * i);
}
template <class F> __global__ void k_sqr_chain(const uint32_t* a, uint32_t* out, int iters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    F x = F::load(a + F::N * (i % 1024));
    for (int it = 0; it < iters; it++) x = x.sqr();
    if (x.v[0] == 0x12345678u) x.store(out + F::N * (i % 1024));
}
template <class F> __global__ void k_ops(const uint32_t* a, const uint32_t* b, uint32_t* mul, uint32_t* add, uint32_t* sub, uint32_t* inv, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    F x = F::load(a + F::N * i), y = F::load(b + F::N * i);
    (x * y).store(mul + F::N * i); (x + y).store(add + F::N * i); (x - y).store(sub + F::N * i);
    (x.inverse() * x).store(inv + F::N * i);
}
template <class F, int ILP> __global__ void k_chain(const uint32_t* a, uint32_t* out, int iters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    F x[ILP]; F y = F::load(a + F::N * (i % 1024));
#pragma unroll
    for (int k = 0; k < ILP; k++) { x[k] = F::load(a + F::N * ((i + k * 7) % 1024)); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) x[k] = x[k] * y;
    }
    F acc = x[0];
#pragma unroll
    for (int k = 1; k < ILP; k++) acc = acc + x[k];
    if (acc.v[0] == 0x12345678u) acc.store(out + F::N * (i % 1024));
}

template <class F, class P, int N64> static int check(const char* name, uint64_t inv64) {
    const int n = 4096;
    uint64_t m[N64]; for (int i = 0; i < N64; i++) m[i] = (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32);
    std::vector<uint64_t> a(n * N64), b(n * N64);
    srand(7);
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < N64; j++) { a[i * N64 + j] = ((uint64_t)rand() << 42) ^ ((uint64_t)rand() << 21) ^ rand(); b[i * N64 + j] = ((uint64_t)rand() << 42) ^ ((uint64_t)rand() << 21) ^ rand(); }
        a[i * N64 + N64 - 1] %= m[N64 - 1]; b[i * N64 + N64 - 1] %= m[N64 - 1];   // < p
    }
    // edge values
    memset(&a[0], 0, 8 * N64); for (int j = 0; j < N64; j++) a[N64 + j] = m[j]; a[N64] -= 1;   // 0, p-1
    memcpy(&b[0], &a[N64], 8 * N64); memcpy(&b[N64], &a[N64], 8 * N64);
    uint32_t *da, *db, *dm, *dadd, *dsub, *dinv; size_t bytes = n * N64 * 8;
    cudaMalloc(&da, bytes); cudaMalloc(&db, bytes); cudaMalloc(&dm, bytes); cudaMalloc(&dadd, bytes); cudaMalloc(&dsub, bytes); cudaMalloc(&dinv, bytes);
    cudaMemcpy(da, a.data(), bytes, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), bytes, cudaMemcpyHostToDevice);
    k_ops<F><<<n / 128, 128>>>(da, db, dm, dadd, dsub, dinv, n);
    std::vector<uint64_t> rm(n * N64), ra(n * N64), rs(n * N64), ri(n * N64);
    cudaMemcpy(rm.data(), dm, bytes, cudaMemcpyDeviceToHost); cudaMemcpy(ra.data(), dadd, bytes, cudaMemcpyDeviceToHost);
    cudaMemcpy(rs.data(), dsub, bytes, cudaMemcpyDeviceToHost); cudaMemcpy(ri.data(), dinv, bytes, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return 1; }
    int bad = 0;
    uint64_t one[N64]; for (int i = 0; i < N64; i++) one[i] = (uint64_t)P::r1(2 * i) | ((uint64_t)P::r1(2 * i + 1) << 32);
    for (int i = 0; i < n; i++) {
        uint64_t e1[N64]; host_mont_mul<N64>(e1, &a[i * N64], &b[i * N64], m, inv64);
        if (memcmp(e1, &rm[i * N64], 8 * N64)) { if (bad < 5) printf("%s mul mismatch at %d\n", name, i); bad++; }
        // add/sub reference
        uint64_t s[N64], d[N64]; u128 c = 0; for (int j = 0; j < N64; j++) { c += (u128)a[i * N64 + j] + b[i * N64 + j]; s[j] = (uint64_t)c; c >>= 64; }
        bool ge = true; for (int j = N64 - 1; j >= 0; j--) if (s[j] != m[j]) { ge = s[j] > m[j]; break; }
        if (ge) { uint64_t br = 0; for (int j = 0; j < N64; j++) { u128 t = (u128)s[j] - m[j] - br; s[j] = (uint64_t)t; br = (uint64_t)(t >> 64) & 1; } }
        uint64_t br = 0; for (int j = 0; j < N64; j++) { u128 t = (u128)a[i * N64 + j] - b[i * N64 + j] - br; d[j] = (uint64_t)t; br = (uint64_t)(t >> 64) & 1; }
        if (br) { c = 0; for (int j = 0; j < N64; j++) { c += (u128)d[j] + m[j]; d[j] = (uint64_t)c; c >>= 64; } }
        if (memcmp(s, &ra[i * N64], 8 * N64)) { if (bad < 5) printf("%s add mismatch at %d\n", name, i); bad++; }
        if (memcmp(d, &rs[i * N64], 8 * N64)) { if (bad < 5) printf("%s sub mismatch at %d\n", name, i); bad++; }
        bool az = true; for (int j = 0; j < N64; j++) az &= a[i * N64 + j] == 0;
        if (!az && memcmp(one, &ri[i * N64], 8 * N64)) { if (bad < 5) printf("%s inverse mismatch at %d\n", name, i); bad++; }
    }
    k_sqr_check<F><<<n / 128, 128>>>(da, dm, n);
    cudaMemcpy(rm.data(), dm, bytes, cudaMemcpyDeviceToHost);
    for (int i = 0; i < n * N64; i++) if (rm[i] != 0) { if (bad < 5) printf("%s sqr mismatch at %d\n", name, i / N64); bad++; break; }
    printf("%s: %d checks, %d mismatches\n", name, n * 5, bad);
    // throughput
    uint32_t* dout; cudaMalloc(&dout, 1024 * N64 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    for (int tpb : {128, 256, 512}) for (int bps : {1, 2, 4}) {
        int iters = 2000; int blocks = sm * bps; float ms;
        k_chain<F, 1><<<blocks, tpb>>>(da, dout, 10); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_chain<F, 1><<<blocks, tpb>>>(da, dout, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        double r1 = (double)blocks * tpb * iters / (ms * 1e-3);
        k_chain<F, 2><<<blocks, tpb>>>(da, dout, 10); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_chain<F, 2><<<blocks, tpb>>>(da, dout, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        double r2 = (double)blocks * tpb * iters * 2 / (ms * 1e-3);
        cudaEventRecord(e0); k_sqr_chain<F><<<blocks, tpb>>>(da, dout, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        double rs = (double)blocks * tpb * iters / (ms * 1e-3);
        printf("%s mul/s: tpb=%d blocks/SM=%d  ILP1 %.3e  ILP2 %.3e  sqr %.3e\n", name, tpb, bps, r1, r2, rs);
    }
    return bad;
}
int main() {
    int bad = 0;
    bad += check<Fr, FrParams, 4>("Fr", 725501752471715839ull);
    bad += check<Fq, FqParams, 6>("Fq", 9586122913090633727ull);
    printf(bad ? "FAIL\n" : "PASS\n");
    return bad ? 1 : 0;
}

// Radix-2 FFT whose coefficients are G1 points and whose twiddles are Fr scalars: the generic
// `EvaluationDomain::{fft,ifft}` with T = G1Projective (algorithms/src/fft/domain.rs:169-221, butterflies :651-664),
// whose one production caller is UniversalParams::lagrange_basis (polycommit/kzg10/data_structures.rs:68-72):
//     basis = domain.ifft(powers_of_beta_g[0..n] as projective);  batch_normalization_into_affine(basis)
// The outputs are affine (canonical), so the result is bit-identical to the reference's whatever ladder and
// coordinate system computes it.  Here: XYZZ coordinates, decimation-in-frequency stages over an n × 192 B array
// in HBM (one thread per butterfly: lo' = lo + hi, hi' = ω^{±e}·(lo − hi) by MSB-first double-and-add over the
// canonical twiddle), then one kernel that applies n^{-1} (inverse only), normalises and stores to the bit-reversed
// address so that the output is in natural order.  The twiddles are the NTT's cached table (ω^{-e} = −ω^{n/2−e}
// becomes ω^{n/2−e}·(hi − lo)); the last stage has e = 0 and multiplies nothing.
// Cost: (lg − 1)·n/2 + n scalar multiplications of ≈ 4·10^3 Fq mul each — compute-bound, ≈ 0.1 s at n = 2^16.
#include "poly.cuh"

#include "ec.cuh"
#include "ff.cuh"
#include "msm.cuh"   // count_launch, ensure_pool_configured

namespace b200 {

// k·P for a Montgomery-form scalar: MSB-first double-and-add over k.to_bigint()
__device__ __noinline__ XYZZ xyzz_mul_fr(const XYZZ& P, const Fr& k_mont) {
    const Fr k = k_mont.from_mont();
    XYZZ acc = XYZZ::infinity();
    if (P.is_inf()) return acc;
    bool started = false;
#pragma unroll 1
    for (int bit = 255; bit >= 0; bit--) {
        const bool b = (k.v[bit >> 5] >> (bit & 31)) & 1u;
        if (started) acc.dbl();
        if (b) { acc.add(P); started = true; }
    }
    return acc;
}

__global__ void k_g1_to_xyzz(const uint8_t* __restrict__ pts, size_t stride, size_t n, uint32_t* __restrict__ X) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    XYZZ::from_affine(load_affine(pts, stride, i)).store(X + i * XYZZ_WORDS);
}

// one DIF stage with gap = 2^s
__global__ void __launch_bounds__(32) k_g1_ntt_stage(uint32_t* __restrict__ X, int lg, int s, const Fr* __restrict__ tw, int lgN, int inverse) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x, half = (size_t)1 << (lg - 1);
    if (b >= half) return;
    const size_t gap = (size_t)1 << s, k = b & (gap - 1), ch = b >> s;
    uint32_t* plo = X + (ch * 2 * gap + k) * XYZZ_WORDS;
    uint32_t* phi = plo + gap * XYZZ_WORDS;
    XYZZ lo = XYZZ::load(plo), hi = XYZZ::load(phi);
    XYZZ sum = lo;
    sum.add(hi);
    sum.store(plo);
    const size_t e = k << (lg - 1 - s);                  // twiddle exponent: ω_n^{±e}, e < n/2
    XYZZ d;
    if (inverse && e != 0) { d = hi; lo.Y = lo.Y.neg(); d.add(lo); }      // hi − lo, to be scaled by ω^{n/2−e} = −ω^{−e}
    else { d = lo; hi.Y = hi.Y.neg(); d.add(hi); }                         // lo − hi
    if (e != 0) {
        const size_t idx = (inverse ? half - e : e) << (lgN - lg);
        d = xyzz_mul_fr(d, tw[idx]);
    }
    d.store(phi);
}

// out[bitrev(i)] = affine(scale · X[i])
__global__ void __launch_bounds__(32) k_g1_ntt_finish(const uint32_t* __restrict__ X, int lg, int inverse, uint8_t* __restrict__ out, size_t stride) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)1 << lg;
    if (i >= n) return;
    XYZZ v = XYZZ::load(X + i * XYZZ_WORDS);
    if (inverse && lg > 0) {
        Fr h = Fr::one();
        for (int k = 0; k < lg; k++) h = h.half();       // size_inv = 2^{-lg} (domain.rs:138-139)
        v = xyzz_mul_fr(v, h);
    }
    const size_t r = lg ? (size_t)(__brevll((unsigned long long)i) >> (64 - lg)) : 0;
    store_affine(out, stride, r, v.to_affine());
}

int g1_ntt_device(void* d_out, size_t out_stride, const void* d_in, size_t in_stride, uint32_t lg, int direction, cudaStream_t stream) {
    if (!d_out || !d_in || lg > 26 || in_stride < 104 || (in_stride & 7) || out_stride < 104 || (out_stride & 7) || direction < 0 || direction > 1)
        return (int)cudaErrorInvalidValue;
    const size_t n = (size_t)1 << lg;
    const void* tw = nullptr;
    int lgN = 0, rc = 0;
    if (lg > 0 && (rc = ntt_get_twiddles((int)lg, &tw, &lgN)) != 0) return rc;
    uint32_t* X = nullptr;
    cudaError_t e = pool_alloc(&X, n * XYZZ_WORDS * 4, stream);
    if (e != cudaSuccess) return (int)e;
    k_g1_to_xyzz<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>((const uint8_t*)d_in, in_stride, n, X);
    count_launch();
    for (int s = (int)lg - 1; s >= 0; s--) {
        k_g1_ntt_stage<<<(unsigned)((n / 2 + 31) / 32), 32, 0, stream>>>(X, (int)lg, s, (const Fr*)tw, lgN, direction);
        count_launch();
    }
    k_g1_ntt_finish<<<(unsigned)((n + 31) / 32), 32, 0, stream>>>(X, (int)lg, direction, (uint8_t*)d_out, out_stride);
    count_launch();
    rc = (int)cudaGetLastError();
    cudaFreeAsync(X, stream);
    return rc;
}

}  // namespace b200

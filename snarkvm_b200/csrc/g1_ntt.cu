// Radix-2 FFT whose coefficients are G1 points and whose twiddles are Fr scalars: the generic
// `EvaluationDomain::{fft,ifft}` with T = G1Projective (algorithms/src/fft/domain.rs:169-221, butterflies :651-664),
// whose one production caller is UniversalParams::lagrange_basis (polycommit/kzg10/data_structures.rs:68-72):
//     basis = domain.ifft(powers_of_beta_g[0..n] as projective);  batch_normalization_into_affine(basis)
// The outputs are affine (canonical), so the result is bit-identical to the reference's whatever ladder and
// coordinate system computes it.
//
// Structure: an n × 192 B XYZZ array in HBM, one decimation-in-frequency stage per launch, one thread per butterfly
// (lo' = lo + hi, hi' = ω^{±e}·(lo − hi)), then one kernel that normalises and stores to the bit-reversed address so
// that the output is in natural order.  The work is (lg − 1)·n/2 + n/2 scalar multiplications by 253-bit twiddles —
// compute-bound by three orders of magnitude — so everything goes into the scalar multiplication:
//   * signed radix-8 digits (−4 … 4), the SAME window boundaries in every lane, so a warp never executes an addition
//     for a minority of its lanes (a NAF would: its non-zero digits sit at lane-dependent positions);
//   * the table {d, 2d, 3d, 4d} is made AFFINE with one field inversion shared by the CTA (cta_inverse.cuh), so the
//     85 window additions are mixed additions (8M + 2S instead of 12M + 2S) and the table is 384 B per thread:
//     48 KB of shared memory per 128-thread CTA, four CTAs per SM;
//   * 255 doublings + 85 mixed additions ≈ 3.2·10^3 Fq multiplications per twiddle (double-and-add: 4.0·10^3, and half
//     of its additions would run for a few lanes only);
//   * inverse transforms fold n^{-1} into the first stage (its twiddles become n^{-1}·ω^{-e}, its untwiddled outputs
//     get a plain n^{-1} pass) — n/2 extra scalar multiplications instead of n;
//   * the final normalisation shares one inversion per CTA as well.
// The twiddles are the NTT's cached table (ω^{-e} = −ω^{n/2−e} becomes ω^{n/2−e}·(hi − lo)).
#include "poly.cuh"

#define FF_CALL_MUL 1
#include "ec.cuh"
#include "cta_inverse.cuh"
#include "msm.cuh"   // count_launch, pool_alloc

#include <mutex>

namespace b200 {

struct FrArg8 { uint32_t v[8]; };

static constexpr int G1N_THREADS = CTA_INV_THREADS;                 // 128
static constexpr int G1N_TABLE_U4 = 4 * 6 * G1N_THREADS;            // [entry 4][16-byte chunk 6][thread]: 48 KB
static constexpr int G1N_SMEM = G1N_TABLE_U4 * 16;                  // the shared inversion (6 KB) runs inside it before the table is written

// k·d for a canonical (non-Montgomery) 253-bit k.  All 128 threads of the CTA must call it (barriers inside); threads
// with !active return d unchanged.  smem: G1N_SMEM bytes.
static __device__ __noinline__ XYZZ xyzz_mul_windowed(const XYZZ& d, const uint32_t (&k)[8], bool active, uint4* smem) {
    const int tid = threadIdx.x;
    const bool live = active && !d.is_inf();
    // ---- signed radix-8 digits, least significant first: 85 windows × 4 bits (sign | magnitude ≤ 4) in 11 words ----
    uint32_t dig[11];
#pragma unroll
    for (int i = 0; i < 11; i++) dig[i] = 0u;
    {
        uint32_t carry = 0;
#pragma unroll
        for (int w = 0; w < 85; w++) {
            const int bit = 3 * w, wi = bit >> 5, sh = bit & 31;
            const uint32_t lo = k[wi], hi = wi + 1 < 8 ? k[wi + 1] : 0u;
            const uint32_t raw = (__funnelshift_r(lo, hi, sh) & 7u) + carry;
            const uint32_t neg = raw > 4u ? 1u : 0u;
            const uint32_t mag = neg ? 8u - raw : raw;
            carry = neg;
            dig[w >> 3] |= ((neg << 3) | mag) << (4 * (w & 7));
        }
    }
    // ---- table d, 2d, 3d, 4d → affine through one shared inversion ----
    XYZZ t1 = d, t2 = d, t3, t4;
    t2.dbl();
    t3 = t2; t3.add(t1);
    t4 = t2; t4.dbl();
    Fq z1 = t1.ZZ * t1.ZZZ, z2 = t2.ZZ * t2.ZZZ, z3 = t3.ZZ * t3.ZZZ, z4 = t4.ZZ * t4.ZZZ;
    // a multiple that is ∞ (d outside the prime-order subgroup: order 2, 3 or 4) must not zero the CTA's shared product:
    // it contributes 1 and its table entry becomes (0, 0), which the ladder reads as "add nothing"
    const bool f2 = live && !z2.is_zero(), f3 = live && !z3.is_zero(), f4 = live && !z4.is_zero();
    if (!live) z1 = Fq::one();
    if (!f2) z2 = Fq::one();
    if (!f3) z3 = Fq::one();
    if (!f4) z4 = Fq::one();
    Fq p2 = z1 * z2, p3 = p2 * z3, p4 = p3 * z4;
    Fq inv = cta_shared_inverse_by(p4, reinterpret_cast<uint32_t*>(smem), (int)(blockIdx.x & 3u));
    __syncthreads();                                                  // everyone has read its inverse: the table may overwrite the area
    {
        Fq i4 = inv * p3; inv = inv * z4;
        Fq i3 = inv * p2; inv = inv * z3;
        Fq i2 = inv * z1; Fq i1 = inv * z2;
        auto put = [&](int e, const XYZZ& t, const Fq& iz, bool finite) {
            Fq x = t.X * (iz * t.ZZZ), y = t.Y * (iz * t.ZZ);
            if (!finite) { x = Fq::zero(); y = Fq::zero(); }
            uint4* s = smem + (size_t)e * 6 * G1N_THREADS + tid;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                s[c * G1N_THREADS] = make_uint4(x.v[4 * c], x.v[4 * c + 1], x.v[4 * c + 2], x.v[4 * c + 3]);
                s[(3 + c) * G1N_THREADS] = make_uint4(y.v[4 * c], y.v[4 * c + 1], y.v[4 * c + 2], y.v[4 * c + 3]);
            }
        };
        put(0, t1, i1, live); put(1, t2, i2, f2); put(2, t3, i3, f3); put(3, t4, i4, f4);
    }
    // (each thread reads back only its own column of the table: no barrier needed)
    XYZZ acc = XYZZ::infinity();
#pragma unroll 1
    for (int w = 84; w >= 0; w--) {
        acc.dbl(); acc.dbl(); acc.dbl();
        uint32_t word = 0u;
#pragma unroll
        for (int i = 0; i < 11; i++) if (i == (w >> 3)) word = dig[i];
        const uint32_t dg = (word >> (4 * (w & 7))) & 15u, mag = dg & 7u;
        if (live && mag != 0u) {
            const uint4* s = smem + (size_t)(mag - 1u) * 6 * G1N_THREADS + tid;
            AffinePoint q;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                uint4 a = s[c * G1N_THREADS], b = s[(3 + c) * G1N_THREADS];
                q.x.v[4 * c] = a.x; q.x.v[4 * c + 1] = a.y; q.x.v[4 * c + 2] = a.z; q.x.v[4 * c + 3] = a.w;
                q.y.v[4 * c] = b.x; q.y.v[4 * c + 1] = b.y; q.y.v[4 * c + 2] = b.z; q.y.v[4 * c + 3] = b.w;
            }
            q.inf = q.x.is_zero() && q.y.is_zero();
            acc.add_affine(q, (dg >> 3) != 0u);
        }
    }
    __syncthreads();                                                  // the next call reuses the shared memory
    return live ? acc : d;
}

__global__ void k_g1_to_xyzz(const uint8_t* __restrict__ pts, size_t stride, size_t n, uint32_t* __restrict__ X) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    XYZZ::from_affine(load_affine(pts, stride, i)).store(X + i * XYZZ_WORDS);
}

// MODE 0: one DIF stage with gap = 2^s: lo' = lo + hi, hi' = (scale·)ω^{±e}·(lo − hi).
// MODE 1: the untwiddled outputs of that stage (the lo halves) times `scale` — the n^{-1} of an inverse transform,
//         applied once, right after its first stage.
template <int MODE>
__global__ void __launch_bounds__(G1N_THREADS, 4) k_g1_ntt_stage(uint32_t* __restrict__ X, int lg, int s, const Fr* __restrict__ tw, int lgN,
                                                                 int inverse, int with_scale, FrArg8 scale_mont) {
    extern __shared__ uint4 g1n_smem[];
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x, half = (size_t)1 << (lg - 1);
    const bool in_range = b < half;
    const size_t gap = (size_t)1 << s, k = b & (gap - 1), ch = b >> s;
    uint32_t* plo = X + (ch * 2 * gap + k) * XYZZ_WORDS;
    uint32_t* phi = plo + gap * XYZZ_WORDS;
    Fr scale;
#pragma unroll
    for (int i = 0; i < 8; i++) scale.v[i] = scale_mont.v[i];
    XYZZ d = XYZZ::infinity();
    Fr kf = Fr::one();
    bool do_mul = false;
    if (MODE == 1) {
        if (in_range) { d = XYZZ::load(plo); kf = scale; do_mul = true; }
    } else if (in_range) {
        XYZZ lo = XYZZ::load(plo), hi = XYZZ::load(phi);
        XYZZ sum = lo;
        sum.add(hi);
        sum.store(plo);
        const size_t e = k << (lg - 1 - s);                  // twiddle exponent: ω_n^{±e}, e < n/2
        if (inverse && e != 0) { d = hi; lo.Y = lo.Y.neg(); d.add(lo); }      // hi − lo, to be scaled by ω^{n/2−e} = −ω^{−e}
        else { d = lo; hi.Y = hi.Y.neg(); d.add(hi); }                         // lo − hi
        if (e != 0) {
            kf = tw[(inverse ? half - e : e) << (lgN - lg)];
            if (with_scale) kf = kf * scale;
            do_mul = true;
        } else if (with_scale) { kf = scale; do_mul = true; }
    }
    const Fr kc = kf.from_mont();
    uint32_t kw[8];
#pragma unroll
    for (int i = 0; i < 8; i++) kw[i] = kc.v[i];
    // stages whose twiddles are all 1 (the last one) skip the ladder for the whole CTA
    if (__syncthreads_or(do_mul ? 1 : 0)) d = xyzz_mul_windowed(d, kw, do_mul, g1n_smem);
    if (in_range) d.store(MODE == 1 ? plo : phi);
}

// out[bitrev(i)] = affine(X[i]) with one inversion per CTA
__global__ void __launch_bounds__(G1N_THREADS, 4) k_g1_ntt_finish(const uint32_t* __restrict__ X, int lg, uint8_t* __restrict__ out, size_t stride) {
    __shared__ uint4 sh[CTA_INV_SMEM_BYTES / 16];
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)1 << lg;
    XYZZ v = XYZZ::infinity();
    if (i < n) v = XYZZ::load(X + i * XYZZ_WORDS);
    const bool inf = v.is_inf();
    Fq z = inf ? Fq::one() : v.ZZ * v.ZZZ;
    Fq iz = cta_shared_inverse_by(z, reinterpret_cast<uint32_t*>(sh), (int)(blockIdx.x & 3u));
    if (i >= n) return;
    AffinePoint a;
    if (inf) { a.x = Fq::zero(); a.y = Fq::one(); a.inf = true; }   // Affine::zero(), affine.rs:57-59
    else { a.x = v.X * (iz * v.ZZZ); a.y = v.Y * (iz * v.ZZ); a.inf = false; }
    const size_t r = lg ? (size_t)(__brevll((unsigned long long)i) >> (64 - lg)) : 0;
    store_affine(out, stride, r, a);
}

__global__ void k_fr_size_inv(Fr* out, int lg) {
    Fr h = Fr::one();
    for (int k = 0; k < lg; k++) h = h.half();           // size_inv = 2^{-lg} (domain.rs:138-139)
    *out = h;
}

int g1_ntt_device(void* d_out, size_t out_stride, const void* d_in, size_t in_stride, uint32_t lg, int direction, cudaStream_t stream) {
    if (!d_out || !d_in || lg > 26 || in_stride < 104 || (in_stride & 7) || out_stride < 104 || (out_stride & 7) || direction < 0 || direction > 1)
        return (int)cudaErrorInvalidValue;
    const size_t n = (size_t)1 << lg;
    const void* tw = nullptr;
    int lgN = 0, rc = 0;
    if (lg > 0 && (rc = ntt_get_twiddles((int)lg, &tw, &lgN)) != 0) return rc;
    {
        static std::once_flag once[64];
        int dev = 0; cudaGetDevice(&dev);
        std::call_once(once[dev & 63], [] {
            cudaFuncSetAttribute(k_g1_ntt_stage<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, G1N_SMEM);
            cudaFuncSetAttribute(k_g1_ntt_stage<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, G1N_SMEM);
        });
    }
    uint32_t* X = nullptr;
    cudaError_t e = pool_alloc(&X, n * XYZZ_WORDS * 4 + 256, stream);
    if (e != cudaSuccess) return (int)e;
    k_g1_to_xyzz<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>((const uint8_t*)d_in, in_stride, n, X);
    count_launch();
    FrArg8 scale = {};
    if (direction == 1 && lg > 0) {
        Fr* d_scale = (Fr*)(X + n * XYZZ_WORDS);
        k_fr_size_inv<<<1, 1, 0, stream>>>(d_scale, (int)lg);
        count_launch();
        if ((rc = (int)cudaMemcpyAsync(scale.v, d_scale, 32, cudaMemcpyDeviceToHost, stream)) != 0 ||
            (rc = (int)cudaStreamSynchronize(stream)) != 0) { cudaFreeAsync(X, stream); return rc; }
    }
    const unsigned grid = (unsigned)((n / 2 + G1N_THREADS - 1) / G1N_THREADS);
    for (int s = (int)lg - 1; s >= 0; s--) {
        const int with_scale = (direction == 1 && s == (int)lg - 1) ? 1 : 0;
        k_g1_ntt_stage<0><<<grid, G1N_THREADS, G1N_SMEM, stream>>>(X, (int)lg, s, (const Fr*)tw, lgN, direction, with_scale, scale);
        count_launch();
        if (with_scale) {
            k_g1_ntt_stage<1><<<grid, G1N_THREADS, G1N_SMEM, stream>>>(X, (int)lg, s, (const Fr*)tw, lgN, direction, 1, scale);
            count_launch();
        }
    }
    k_g1_ntt_finish<<<(unsigned)((n + G1N_THREADS - 1) / G1N_THREADS), G1N_THREADS, 0, stream>>>(X, (int)lg, (uint8_t*)d_out, out_stride);
    count_launch();
    rc = (int)cudaGetLastError();
    cudaFreeAsync(X, stream);
    return rc;
}

}  // namespace b200

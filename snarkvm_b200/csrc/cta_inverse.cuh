// One field inversion shared by a 128-thread CTA (Montgomery's trick across threads).
//
// The reference batches inversions the same way on the CPU (batch_inversion, fields/src/lib.rs:78-129; used by
// batch_add in msm/variable_base/batched.rs:175-325 and batch_normalization in the projective templates): multiply the
// values together, invert the product once, peel the individual inverses off.  Here the batch is the CTA: every thread
// contributes one non-zero Fq, ONE warp multiplies the 128 values (4 per lane, then a shuffle scan across lanes), inverts the
// single total with the limb-per-lane Fermat ladder (coop_inverse, ff.cuh) and unwinds.  In SIMT time an inversion costs a warp
// the same whether 1 or 32 lanes need it, so this is ≈ 4× cheaper than one inversion per thread even before the cooperative
// ladder.  Include after ec.cuh / ff.cuh; every function is static to its translation unit.
#pragma once
#include "ff.cuh"

namespace b200 {

static constexpr int CTA_INV_THREADS = 128;
static constexpr int CTA_INV_SMEM_BYTES = CTA_INV_THREADS * 48;
FF_DEV Fq shfl_up_fq(const Fq& a, int d) {
    Fq r;
#pragma unroll
    for (int j = 0; j < 12; j++) r.v[j] = __shfl_up_sync(0xffffffffu, a.v[j], d);
    return r;
}
FF_DEV Fq shfl_down_fq(const Fq& a, int d) {
    Fq r;
#pragma unroll
    for (int j = 0; j < 12; j++) r.v[j] = __shfl_down_sync(0xffffffffu, a.v[j], d);
    return r;
}
FF_DEV Fq shfl_idx_fq(const Fq& a, int l) {
    Fq r;
#pragma unroll
    for (int j = 0; j < 12; j++) r.v[j] = __shfl_sync(0xffffffffu, a.v[j], l);
    return r;
}
// One Fermat inversion per CTA instead of one per warp-lane: the 128 running products (all non-zero) go through shared memory,
// ONE warp multiplies them together (4 per lane, then a shuffle scan across lanes), inverts the total and unwinds.  In SIMT
// time an inversion costs a warp ≈ 515 Fq mul whether 1 or 32 lanes need it, so the per-thread version spends 4 × 515 per
// CTA and this one ≈ 540.  The inverting warp rotates with blockIdx so co-resident CTAs load different SM sub-partitions.
static __device__ __noinline__ Fq cta_shared_inverse(const Fq& run, uint32_t* sh) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    run.store(sh + tid * 12);
    __syncthreads();
    if (warp == (int)(blockIdx.x & 3u)) {
        uint32_t* mine = sh + lane * 48;
        Fq a0 = Fq::load(mine), a1 = Fq::load(mine + 12), a2 = Fq::load(mine + 24), a3 = Fq::load(mine + 36);
        Fq p1 = a0 * a1, p2 = p1 * a2, p3 = p2 * a3;
        Fq incl = p3, suff = p3;                           // inclusive prefix / suffix products over lanes
#pragma unroll 1
        for (int d = 1; d < 32; d <<= 1) {
            Fq up = shfl_up_fq(incl, d), dn = shfl_down_fq(suff, d);
            if (lane >= d) incl = incl * up;
            if (lane + d < 32) suff = suff * dn;
        }
        Fq tinv = shfl_idx_fq(incl, 31).inverse();
        Fq before = shfl_up_fq(incl, 1), after = shfl_down_fq(suff, 1);
        Fq ip3 = tinv;                                     // 1 / p3 of this lane = tinv · Π(other lanes)
        if (lane > 0) ip3 = ip3 * before;
        if (lane < 31) ip3 = ip3 * after;
        Fq ip2 = ip3 * a3, ip1 = ip2 * a2;
        (ip3 * p2).store(mine + 36);                       // 1/a3
        (ip2 * p1).store(mine + 24);                       // 1/a2
        (ip1 * a0).store(mine + 12);                       // 1/a1
        (ip1 * a1).store(mine);                            // 1/a0
    }
    __syncthreads();
    return Fq::load(sh + tid * 12);
}

// One Fermat inversion per CTA (see cta_shared_inverse above), here with the inverting warp chosen by the caller.
static __device__ __noinline__ Fq cta_shared_inverse_by(const Fq& run, uint32_t* sh, int inv_warp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    run.store(sh + tid * 12);
    __syncthreads();
    if (warp == inv_warp) {
        uint32_t* mine = sh + lane * 48;
        Fq a0 = Fq::load(mine), a1 = Fq::load(mine + 12), a2 = Fq::load(mine + 24), a3 = Fq::load(mine + 36);
        Fq p1 = a0 * a1, p2 = p1 * a2, p3 = p2 * a3;
        Fq incl = p3, suff = p3;                           // inclusive prefix / suffix products over lanes
#pragma unroll 1
        for (int d = 1; d < 32; d <<= 1) {
            Fq up = shfl_up_fq(incl, d), dn = shfl_down_fq(suff, d);
            if (lane >= d) incl = incl * up;
            if (lane + d < 32) suff = suff * dn;
        }
        Fq tinv = coop_inverse<FqParams>(shfl_idx_fq(incl, 31));   // one element, limb-per-lane: ≈ 4× fewer instructions than 32 redundant chains
        Fq before = shfl_up_fq(incl, 1), after = shfl_down_fq(suff, 1);
        Fq ip3 = tinv;                                     // 1 / p3 of this lane = tinv · Π(other lanes)
        if (lane > 0) ip3 = ip3 * before;
        if (lane < 31) ip3 = ip3 * after;
        Fq ip2 = ip3 * a3, ip1 = ip2 * a2;
        (ip3 * p2).store(mine + 36);                       // 1/a3
        (ip2 * p1).store(mine + 24);                       // 1/a2
        (ip1 * a0).store(mine + 12);                       // 1/a1
        (ip1 * a1).store(mine);                            // 1/a0
    }
    __syncthreads();
    return Fq::load(sh + tid * 12);
}


}  // namespace b200

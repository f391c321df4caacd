// Four lanes per group operation — the latency path of small MSMs.
//
// A lone warp needs ≈ 1.0 µs per Fq multiplication (profiles/ff_microbench_r1.log: one warp per scheduler runs the
// multiplier at 69 % of its saturated rate, two independent chains per thread gain 2 %), so the tail of a small MSM — a
// few dozen DEPENDENT point additions of 14 multiplications each (bucket running sums, window combine) — is ≈ 18 µs per
// addition whatever the occupancy.  Lanes are free there: the whole tail is a few hundred points.  Here the four lanes
// 4k … 4k+3 of a warp (a "quad") hold identical copies of the operands of ONE group operation and each lane computes a
// different one of its independent field multiplications; the products are exchanged with width-4 shuffles.  An XYZZ
// addition becomes 4 multiplication steps instead of 14, a doubling 3 instead of 9, a mixed addition 4 instead of 10.
//
// Invariant: every lane of a quad holds bit-identical operands and ends with bit-identical results (all lanes execute the
// same instructions on the same data apart from the operand selection of each step).  Quads of one warp may diverge from
// each other (∞ operands, equal points): the exchanges name only the quad's own lanes in their mask.
#pragma once
#include "ec.cuh"

namespace b200 {

struct Quad {
    int q;             // lane within the quad (0 … 3)
    int slot;          // quad within the warp (0 … 7)
    uint32_t mask;     // the quad's four lanes
    FF_DEV static Quad here() {
        Quad r; const uint32_t lane = threadIdx.x & 31u;
        r.q = (int)(lane & 3u); r.slot = (int)(lane >> 2); r.mask = 0xFu << (lane & 28u);
        return r;
    }
};

FF_DEV Fq quad_get(const Fq& v, int src, uint32_t mask) {
    Fq r;
#pragma unroll
    for (int j = 0; j < 12; j++) r.v[j] = __shfl_sync(mask, v.v[j], src, 4);
    return r;
}
FF_DEV Fq quad_sel(int q, const Fq& a0, const Fq& a1, const Fq& a2, const Fq& a3) {
    Fq r;
#pragma unroll
    for (int j = 0; j < 12; j++) {
        const uint32_t lo = (q & 1) ? a1.v[j] : a0.v[j], hi = (q & 1) ? a3.v[j] : a2.v[j];
        r.v[j] = (q & 2) ? hi : lo;
    }
    return r;
}

// dbl-2008-s-1 in three multiplication steps
FF_DEV XYZZ quad_dbl(const XYZZ& a, const Quad& Q) {
    if (a.is_inf()) return a;
    const Fq U = a.Y.dbl();
    // step 1: V = U², XX = X²
    const Fq r1 = quad_sel(Q.q, U, a.X, U, a.X) * quad_sel(Q.q, U, a.X, U, a.X);
    const Fq V = quad_get(r1, 0, Q.mask), XX = quad_get(r1, 1, Q.mask);
    const Fq M = XX.dbl() + XX;
    // step 2: W = U·V, S = X·V, MM = M², ZZ3 = V·ZZ
    const Fq r2 = quad_sel(Q.q, U, a.X, M, V) * quad_sel(Q.q, V, V, M, a.ZZ);
    const Fq W = quad_get(r2, 0, Q.mask), S = quad_get(r2, 1, Q.mask), MM = quad_get(r2, 2, Q.mask);
    XYZZ r;
    r.ZZ = quad_get(r2, 3, Q.mask);
    r.X = MM - S.dbl();
    // step 3: M·(S − X3), W·Y, ZZZ3 = W·ZZZ
    const Fq r3 = quad_sel(Q.q, M, W, W, W) * quad_sel(Q.q, S - r.X, a.Y, a.ZZZ, a.ZZZ);
    r.Y = quad_get(r3, 0, Q.mask) - quad_get(r3, 1, Q.mask);
    r.ZZZ = quad_get(r3, 2, Q.mask);
    return r;
}

// add-2008-s in four multiplication steps
FF_DEV XYZZ quad_add(const XYZZ& a, const XYZZ& b, const Quad& Q) {
    if (b.is_inf()) return a;
    if (a.is_inf()) return b;
    // step 1: U1 = X1·ZZ2, U2 = X2·ZZ1, S1 = Y1·ZZZ2, S2 = Y2·ZZZ1
    const Fq r1 = quad_sel(Q.q, a.X, b.X, a.Y, b.Y) * quad_sel(Q.q, b.ZZ, a.ZZ, b.ZZZ, a.ZZZ);
    const Fq U1 = quad_get(r1, 0, Q.mask), U2 = quad_get(r1, 1, Q.mask), S1 = quad_get(r1, 2, Q.mask), S2 = quad_get(r1, 3, Q.mask);
    const Fq P = U2 - U1, R = S2 - S1;
    if (P.is_zero()) {
        if (R.is_zero()) return quad_dbl(a, Q);
        return XYZZ::infinity();                               // P + (−P)
    }
    // step 2: PP = P², RR = R², ZZ1·ZZ2 (stays in lane 2), ZZZ1·ZZZ2 (stays in lane 3)
    const Fq r2 = quad_sel(Q.q, P, R, a.ZZ, a.ZZZ) * quad_sel(Q.q, P, R, b.ZZ, b.ZZZ);
    const Fq PP = quad_get(r2, 0, Q.mask), RR = quad_get(r2, 1, Q.mask);
    // step 3: PPP = P·PP (lanes 0 and 3), Q = U1·PP, ZZ3 = (ZZ1·ZZ2)·PP
    const Fq r3 = quad_sel(Q.q, P, U1, r2, P) * PP;
    const Fq PPP = quad_get(r3, 0, Q.mask), Qv = quad_get(r3, 1, Q.mask);
    XYZZ r;
    r.ZZ = quad_get(r3, 2, Q.mask);
    r.X = RR - PPP - Qv.dbl();
    // step 4: R·(Q − X3), S1·PPP, ZZZ3 = (ZZZ1·ZZZ2)·PPP in lane 3
    const Fq r4 = quad_sel(Q.q, R, S1, R, r2) * quad_sel(Q.q, Qv - r.X, PPP, PPP, PPP);
    r.Y = quad_get(r4, 0, Q.mask) - quad_get(r4, 1, Q.mask);
    r.ZZZ = quad_get(r4, 3, Q.mask);
    return r;
}

// mixed addition with an affine point whose y may be negated, four multiplication steps
FF_DEV XYZZ quad_add_affine(const XYZZ& a, const AffinePoint& p, bool negate, const Quad& Q) {
    if (p.inf) return a;
    const Fq qy = negate ? p.y.neg() : p.y;
    if (a.is_inf()) { XYZZ r; r.X = p.x; r.Y = qy; r.ZZ = Fq::one(); r.ZZZ = Fq::one(); return r; }
    // step 1: U2 = x·ZZ, S2 = y·ZZZ
    const Fq r1 = quad_sel(Q.q, p.x, qy, p.x, qy) * quad_sel(Q.q, a.ZZ, a.ZZZ, a.ZZ, a.ZZZ);
    const Fq U2 = quad_get(r1, 0, Q.mask), S2 = quad_get(r1, 1, Q.mask);
    const Fq P = U2 - a.X, R = S2 - a.Y;
    if (P.is_zero()) {
        if (!R.is_zero()) return XYZZ::infinity();
        XYZZ d; d.X = p.x; d.Y = qy; d.ZZ = Fq::one(); d.ZZZ = Fq::one();
        return quad_dbl(d, Q);
    }
    // step 2: PP = P², RR = R²
    const Fq r2 = quad_sel(Q.q, P, R, P, R) * quad_sel(Q.q, P, R, P, R);
    const Fq PP = quad_get(r2, 0, Q.mask), RR = quad_get(r2, 1, Q.mask);
    // step 3: PPP = P·PP (lanes 0 and 3), Q = X·PP, ZZ3 = ZZ·PP
    const Fq r3 = quad_sel(Q.q, P, a.X, a.ZZ, P) * PP;
    const Fq PPP = quad_get(r3, 0, Q.mask), Qv = quad_get(r3, 1, Q.mask);
    XYZZ r;
    r.ZZ = quad_get(r3, 2, Q.mask);
    r.X = RR - PPP - Qv.dbl();
    // step 4: R·(Q − X3), Y·PPP, ZZZ3 = ZZZ·PPP
    const Fq r4 = quad_sel(Q.q, R, a.Y, a.ZZZ, a.ZZZ) * quad_sel(Q.q, Qv - r.X, PPP, PPP, PPP);
    r.Y = quad_get(r4, 0, Q.mask) - quad_get(r4, 1, Q.mask);
    r.ZZZ = quad_get(r4, 2, Q.mask);
    return r;
}

// ---- the eight quads ("slots") of a warp; these need the whole warp converged ----
FF_DEV XYZZ slot_shfl_down(const XYZZ& a, int dslots) {
    XYZZ r;
#pragma unroll
    for (int j = 0; j < 12; j++) {
        r.X.v[j] = __shfl_down_sync(0xffffffffu, a.X.v[j], 4 * dslots); r.Y.v[j] = __shfl_down_sync(0xffffffffu, a.Y.v[j], 4 * dslots);
        r.ZZ.v[j] = __shfl_down_sync(0xffffffffu, a.ZZ.v[j], 4 * dslots); r.ZZZ.v[j] = __shfl_down_sync(0xffffffffu, a.ZZZ.v[j], 4 * dslots);
    }
    return r;
}
FF_DEV XYZZ slot_shfl_xor(const XYZZ& a, int mslots) {
    XYZZ r;
#pragma unroll
    for (int j = 0; j < 12; j++) {
        r.X.v[j] = __shfl_xor_sync(0xffffffffu, a.X.v[j], 4 * mslots); r.Y.v[j] = __shfl_xor_sync(0xffffffffu, a.Y.v[j], 4 * mslots);
        r.ZZ.v[j] = __shfl_xor_sync(0xffffffffu, a.ZZ.v[j], 4 * mslots); r.ZZZ.v[j] = __shfl_xor_sync(0xffffffffu, a.ZZZ.v[j], 4 * mslots);
    }
    return r;
}
// every slot ends with Σ over the eight slots (possibly different XYZZ images of the same point in different slots)
FF_DEV XYZZ slot_sum(XYZZ a, const Quad& Q) {
#pragma unroll 1
    for (int m = 4; m >= 1; m >>= 1) {
        __syncwarp();
        const XYZZ o = slot_shfl_xor(a, m);
        a = quad_add(a, o, Q);
    }
    return a;
}
// slot s ends with Σ_{s' ≥ s} a_s'
FF_DEV XYZZ slot_suffix_scan(XYZZ a, const Quad& Q) {
#pragma unroll 1
    for (int d = 1; d < 8; d <<= 1) {
        __syncwarp();
        const XYZZ o = slot_shfl_down(a, d);
        if (Q.slot + d < 8) a = quad_add(a, o, Q);
    }
    return a;
}

}  // namespace b200

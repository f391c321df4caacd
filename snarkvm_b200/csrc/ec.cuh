// BLS12-377 G1 group law on the device (curve y^2 = x^3 + 1, a = 0;
// curves/src/bls12_377/g1.rs:78-91).
//
// Input points use the reference's in-memory `Affine<P>` image
//   { x: Fq, y: Fq, infinity: bool }  — 104-byte stride, Montgomery Fq
//   (curves/src/templates/short_weierstrass_jacobian/affine.rs:41-46).
// Bucket accumulators use extended Jacobian "XYZZ" coordinates
//   x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2, infinity <=> ZZ == 0
// (EFD shortw/xyzz, a = 0: madd-2008-s 8M+2S, add-2008-s 12M+2S, dbl-2008-s-1).
// The group element computed is the same one the reference's Jacobian formulas
// (projective.rs:222-291, 302-339, 407-468) compute; only the final
// to_affine()-normalised image is compared, as in every reference test
// (algorithms/src/msm/variable_base/mod.rs:90-119).
#pragma once
#include "ff.cuh"

namespace b200 {

struct AffinePoint {
    Fq x, y;
    bool inf;
};

// Reference layout: x[48] y[48] inf[1] pad.  The 104-byte stride is only 8-byte aligned
// (odd indices sit at 8 mod 16), so the gather uses 64-bit loads: 12 × LDG.64 per point.
FF_DEV Fq load_fq_u64(const uint8_t* p) {
    Fq r; const uint2* q = reinterpret_cast<const uint2*>(p);
#pragma unroll
    for (int i = 0; i < 6; i++) { uint2 t = __ldg(q + i); r.v[2 * i] = t.x; r.v[2 * i + 1] = t.y; }
    return r;
}
FF_DEV void store_fq_u64(uint8_t* p, const Fq& a) {
    uint2* q = reinterpret_cast<uint2*>(p);
#pragma unroll
    for (int i = 0; i < 6; i++) q[i] = make_uint2(a.v[2 * i], a.v[2 * i + 1]);
}
FF_DEV AffinePoint load_affine(const uint8_t* base, size_t stride, size_t i) {
    const uint8_t* p = base + i * stride;
    AffinePoint a;
    a.x = load_fq_u64(p);
    a.y = load_fq_u64(p + 48);
    a.inf = __ldg(p + 96) != 0;
    return a;
}
FF_DEV void store_affine(uint8_t* base, size_t stride, size_t i, const AffinePoint& a) {
    uint8_t* p = base + i * stride;
    store_fq_u64(p, a.x);
    store_fq_u64(p + 48, a.y);
    // infinity flag + padding as one 8-byte store (stride is a multiple of 8 ≥ 104)
    *reinterpret_cast<unsigned long long*>(p + 96) = a.inf ? 1ull : 0ull;
}

struct XYZZ {
    Fq X, Y, ZZ, ZZZ;

    FF_DEV static XYZZ infinity() { XYZZ r; r.X = Fq::zero(); r.Y = Fq::zero(); r.ZZ = Fq::zero(); r.ZZZ = Fq::zero(); return r; }
    FF_DEV bool is_inf() const { return ZZ.is_zero(); }

    FF_DEV static XYZZ from_affine(const AffinePoint& p) {
        XYZZ r;
        if (p.inf) return infinity();
        r.X = p.x; r.Y = p.y; r.ZZ = Fq::one(); r.ZZZ = Fq::one();
        return r;
    }

    // dbl-2008-s-1 (a = 0)
    FF_DEV void dbl() {
        if (is_inf()) return;
        Fq U = Y.dbl();
        Fq V = U.sqr();
        Fq W = U * V;
        Fq S = X * V;
        Fq XX = X.sqr();
        Fq M = XX.dbl() + XX;
        Fq X3 = M.sqr() - S.dbl();
        Fq Y3 = M * (S - X3) - W * Y;
        X = X3; Y = Y3;
        ZZ = V * ZZ;
        ZZZ = W * ZZZ;
    }

    // mixed addition with an affine point whose y may be negated (signed-digit buckets).
    FF_DEV void add_affine(const AffinePoint& q, bool negate) {
        if (q.inf) return;
        Fq qy = negate ? q.y.neg() : q.y;
        if (is_inf()) { X = q.x; Y = qy; ZZ = Fq::one(); ZZZ = Fq::one(); return; }
        Fq U2 = q.x * ZZ;
        Fq S2 = qy * ZZZ;
        Fq P = U2 - X;
        Fq R = S2 - Y;
        if (P.is_zero()) {
            if (R.is_zero()) {
                // same point: double the affine operand (mdbl-2008-s-1)
                Fq U = qy.dbl();
                Fq V = U.sqr();
                Fq W = U * V;
                Fq S = q.x * V;
                Fq XX = q.x.sqr();
                Fq M = XX.dbl() + XX;
                X = M.sqr() - S.dbl();
                Y = M * (S - X) - W * qy;
                ZZ = V; ZZZ = W;
            } else {
                *this = infinity();          // P + (-P)
            }
            return;
        }
        Fq PP = P.sqr();
        Fq PPP = P * PP;
        Fq Q = X * PP;
        Fq X3 = R.sqr() - PPP - Q.dbl();
        Y = R * (Q - X3) - Y * PPP;
        X = X3;
        ZZ = ZZ * PP;
        ZZZ = ZZZ * PPP;
    }

    // general addition (add-2008-s)
    FF_DEV void add(const XYZZ& o) {
        if (o.is_inf()) return;
        if (is_inf()) { *this = o; return; }
        Fq U1 = X * o.ZZ;
        Fq U2 = o.X * ZZ;
        Fq S1 = Y * o.ZZZ;
        Fq S2 = o.Y * ZZZ;
        Fq P = U2 - U1;
        Fq R = S2 - S1;
        if (P.is_zero()) {
            if (R.is_zero()) dbl(); else *this = infinity();
            return;
        }
        Fq PP = P.sqr();
        Fq PPP = P * PP;
        Fq Q = U1 * PP;
        Fq X3 = R.sqr() - PPP - Q.dbl();
        Y = R * (Q - X3) - S1 * PPP;
        X = X3;
        ZZ = ZZ * o.ZZ * PP;
        ZZZ = ZZZ * o.ZZZ * PPP;
    }

    // k·P for a small public multiplier (used by the bucket reduction: (lo-1)·running)
    FF_DEV XYZZ mul_u32(uint32_t k) const {
        XYZZ acc = infinity();
        bool started = false;
        for (int b = 31; b >= 0; b--) {
            if (started) acc.dbl();
            if ((k >> b) & 1u) { acc.add(*this); started = true; }
        }
        return acc;
    }

    FF_DEV AffinePoint to_affine() const {
        AffinePoint a;
        if (is_inf()) { a.x = Fq::zero(); a.y = Fq::one(); a.inf = true; return a; }   // Affine::zero(), affine.rs:57-59
        // x = X/ZZ, y = Y/ZZZ with one inversion: i = 1/(ZZ·ZZZ)
        Fq i = (ZZ * ZZZ).inverse();
        a.x = X * (i * ZZZ);
        a.y = Y * (i * ZZ);
        a.inf = false;
        return a;
    }

    FF_DEV static XYZZ load(const uint32_t* p) {
        XYZZ r; r.X = Fq::load(p); r.Y = Fq::load(p + 12); r.ZZ = Fq::load(p + 24); r.ZZZ = Fq::load(p + 36); return r;
    }
    FF_DEV void store(uint32_t* p) const { X.store(p); Y.store(p + 12); ZZ.store(p + 24); ZZZ.store(p + 36); }
};

static constexpr int XYZZ_WORDS = 48;   // 192 bytes

}  // namespace b200

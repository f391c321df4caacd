// Fq in radix 2^28 (14 limbs in 32-bit registers), lazily reduced — the arithmetic of the MSM
// bucket-accumulation kernel.
//
// Why: measured on B200 (profiles/README.md, tools/pipe_microbench.cu) a plain IMAD.WIDE.U32 with a
// 64-bit accumulator issues at the full IMAD rate (one warp instruction per 2 clk per SMSP), while the
// carry-chained forms ptxas emits for saturated 32-bit limbs (IMAD.X + IMAD.HI.U32.X pairs) cost ~2.4x
// more per product.  With 28-bit limbs a column accumulator can absorb all 14 + 14 partial products of
// a Montgomery multiplication (28 · 2^60 < 2^64 even for operands with 30-bit limbs) so the multiplier
// is nothing but independent IMAD.WIDE instructions plus a little carry bookkeeping on the ALU pipe, and
// additions / subtractions need no carry propagation at all.
//
// Representation: value = Σ v[i]·2^{28 i}; an element x is stored as x·R' mod p with R' = 2^392
// ("internal Montgomery form").  Values are NOT canonical: every function documents the limb and value
// bounds it needs and guarantees.  The exact op sequences and bounds are model-checked on Python
// integers in tests/manual/fq28_model.py (column sums < 2^64, limbs < 2^32, borrow-free subtraction).
// Conversion from / to the reference's in-memory form (12 × u32, Montgomery R = 2^384; fp_384.rs:52)
// happens once per base and once per bucket item.
#pragma once
#include "../../snarkvm_b200/csrc/ff.cuh"

namespace b200 {

struct Fq28 {
    static constexpr int N = 14;
    static constexpr uint32_t MASK = 0x0fffffffu;
    uint32_t v[N];
};

#define FQ28_TABLE(name, ...) \
    __host__ __device__ static constexpr uint32_t name(int i) { constexpr uint32_t t[14] = {__VA_ARGS__}; return t[i]; }

struct Fq28C {
    // p in radix 2^28 (p ≡ 1 mod 2^28, so −p^{-1} mod 2^28 = 2^28 − 1 and m = −t mod 2^28)
    FQ28_TABLE(p, 0x0000001u, 0x08c0000u, 0x0000085u, 0x5d44300u, 0x800170bu, 0x2fba094u, 0xf1ef362u, 0x00f5138u, 0xa22d9f3u, 0xa1493b1u, 0xb05c06cu, 0x10eac63u, 0xa4617c5u, 0x0001ae3u)
    // 2^400 mod p : mont(x̃, c_in) = x̃·2^8 = x·2^392  (reference form → internal form)
    FQ28_TABLE(c_in, 0xf67abddu, 0xcdbffffu, 0x0d714f0u, 0x5fed6fbu, 0x1adbe5bu, 0xebc50d7u, 0x49c2b74u, 0xd718fccu, 0x56fe5f2u, 0xd413a69u, 0xa9348beu, 0xf3d01dcu, 0xb9c0b1bu, 0x000070eu)
    // 2^384 mod p : mont(x̂, c_out) = x̂·2^-8 = x·2^384  (internal form → reference form)
    FQ28_TABLE(c_out, 0xfffff68u, 0xcdfffffu, 0xfffb102u, 0x9f837ffu, 0xff25140u, 0xa98a7d3u, 0x59f7db3u, 0x6e7c630u, 0xb4e97b7u, 0x3c84e87u, 0x495bf80u, 0xf49a4cfu, 0x661e2fdu, 0x00008d6u)
    // 1 in internal form (2^392 mod p)
    FQ28_TABLE(one, 0xfff67acu, 0x20fffffu, 0xfb0d727u, 0xe9203ffu, 0x249b0e4u, 0xe172345u, 0x955d771u, 0x2bf89aau, 0xb2833b2u, 0x98e116bu, 0x7dc5c97u, 0x0d43e93u, 0x2e3314bu, 0x00003b4u)
    // borrow-proof multiples K·p: limb j carries an extra 2^OFF (taken from limb j+1), so that
    // a + C − b needs no borrow whenever every limb of b is ≤ 2^OFF − 1 and b < K·p.
    FQ28_TABLE(c2p_28, 0x10000002u, 0x1117ffffu, 0x10000109u, 0x1ba885ffu, 0x10002e15u, 0x15f74128u, 0x1e3de6c3u, 0x101ea270u, 0x1445b3e5u, 0x14292762u, 0x160b80d8u, 0x121d58c6u, 0x148c2f89u, 0x000035c6u)
    FQ28_TABLE(c16p_28, 0x10000010u, 0x18bfffffu, 0x1000084fu, 0x1d442fffu, 0x100170b4u, 0x1fba0947u, 0x11ef3621u, 0x10f5138eu, 0x122d9f2fu, 0x11493b19u, 0x105c06c9u, 0x10eac63au, 0x14617c50u, 0x0001ae39u)
    FQ28_TABLE(c8p_30, 0x40000008u, 0x445ffffcu, 0x40000424u, 0x4ea217fcu, 0x4000b856u, 0x47dd04a0u, 0x48f79b0du, 0x407a89c3u, 0x4116cf94u, 0x40a49d89u, 0x482e0361u, 0x48756319u, 0x4230be24u, 0x0000d719u)
};
#undef FQ28_TABLE

// ---------------------------------------------------------------------------------------------
// Montgomery product, interleaved reduction, sliding window of 64-bit column accumulators.
// Needs: limbs of a, b < 2^30 (SQR: < 2^29.6) and value(a)·value(b) < p·R' (≈ 39000·p²).
// Gives: limbs < 2^28 (top limb < 2^15), value < p·(1 + value(a)·value(b)/(p·R')) < 1.01·p for the
//        operand ranges used in this file.
// ---------------------------------------------------------------------------------------------
template <bool SQR>
FF_DEV Fq28 fq28_mul_impl(const Fq28& a, const Fq28& b) {
    constexpr int N = Fq28::N;
    unsigned long long t[2 * N + 1];
#pragma unroll
    for (int k = 0; k < 2 * N + 1; k++) t[k] = 0ull;
#pragma unroll
    for (int i = 0; i < N; i++) {
        if (SQR) {
            t[2 * i] += (unsigned long long)a.v[i] * a.v[i];
#pragma unroll
            for (int j = i + 1; j < N; j++) t[i + j] += (unsigned long long)a.v[i] * (a.v[j] << 1);
        } else {
#pragma unroll
            for (int j = 0; j < N; j++) t[i + j] += (unsigned long long)a.v[j] * b.v[i];
        }
        const uint32_t m = (0u - (uint32_t)t[i]) & Fq28::MASK;            // −t_i mod 2^28
        const unsigned long long c = (t[i] + m) >> 28;                     // p[0] = 1: column i becomes ≡ 0
#pragma unroll
        for (int j = 1; j < N; j++) t[i + j] += (unsigned long long)Fq28C::p(j) * m;
        t[i + 1] += c;
    }
    Fq28 r;
#pragma unroll
    for (int k = N; k < 2 * N - 1; k++) { r.v[k - N] = (uint32_t)t[k] & Fq28::MASK; t[k + 1] += t[k] >> 28; }
    r.v[N - 1] = (uint32_t)t[2 * N - 1];
    return r;
}
// Real (non-inlined) functions: one copy of each ~12 KB body keeps the accumulation loop inside the
// instruction cache (fully inlined the kernel was ~350 KB of SASS and ran instruction-fetch bound).
// ptxas passes and returns the 14-word structs in registers (no stack traffic; checked in SASS).
#ifndef FQ28_INLINE_MUL
static __device__ __noinline__ Fq28 fq28_mul(Fq28 a, Fq28 b) { return fq28_mul_impl<false>(a, b); }
static __device__ __noinline__ Fq28 fq28_sqr(Fq28 a) { return fq28_mul_impl<true>(a, a); }
#else
FF_DEV Fq28 fq28_mul(const Fq28& a, const Fq28& b) { return fq28_mul_impl<false>(a, b); }
FF_DEV Fq28 fq28_sqr(const Fq28& a) { return fq28_mul_impl<true>(a, a); }
#endif

// limb-wise a + b (no carries): limbs add, values add
FF_DEV Fq28 fq28_add(const Fq28& a, const Fq28& b) {
    Fq28 r;
#pragma unroll
    for (int i = 0; i < Fq28::N; i++) r.v[i] = a.v[i] + b.v[i];
    return r;
}
// a + C − b for a borrow-proof constant C = K·p; needs b's limbs ≤ C's limbs (see Fq28C) and b < K·p
#define FQ28_SUB(name, table)                                                        \
    FF_DEV Fq28 name(const Fq28& a, const Fq28& b) {                                 \
        Fq28 r;                                                                      \
        _Pragma("unroll") for (int i = 0; i < Fq28::N; i++) r.v[i] = a.v[i] + Fq28C::table(i) - b.v[i]; \
        return r;                                                                    \
    }
FQ28_SUB(fq28_sub_2p, c2p_28)      // b: limbs < 2^28, value < 2p (a multiplication result)  → result limbs < La + 2^29
FQ28_SUB(fq28_sub_16p, c16p_28)    // b: limbs < 2^28, value < 16p
FQ28_SUB(fq28_sub_8p_wide, c8p_30) // b: limbs < 2^30, value < 8p
#undef FQ28_SUB
// −b for b with limbs < 2^28, value < 2p   (limbs < 2^29, value ≤ 2p)
FF_DEV Fq28 fq28_neg(const Fq28& b) {
    Fq28 r;
#pragma unroll
    for (int i = 0; i < Fq28::N; i++) r.v[i] = Fq28C::c2p_28(i) - b.v[i];
    return r;
}
// carry propagation: limbs < 2^28 afterwards (top limb keeps the rest); value unchanged
FF_DEV Fq28 fq28_norm(const Fq28& a) {
    Fq28 r;
    uint32_t carry = 0;
#pragma unroll
    for (int i = 0; i < Fq28::N - 1; i++) { uint32_t s = a.v[i] + carry; r.v[i] = s & Fq28::MASK; carry = s >> 28; }
    r.v[Fq28::N - 1] = a.v[Fq28::N - 1] + carry;
    return r;
}
FF_DEV Fq28 fq28_const_one() { Fq28 r;
#pragma unroll
    for (int i = 0; i < Fq28::N; i++) r.v[i] = Fq28C::one(i); return r; }

// a ≥ K·p ?  then a −= K·p   (a normalised).  Used only on slow paths / conversions.
FF_DEV void fq28_cond_sub_kp(Fq28& a, uint32_t K) {
    // K·p limbs (normalised, raw top)
    uint32_t kp[Fq28::N];
    unsigned long long carry = 0;
#pragma unroll
    for (int i = 0; i < Fq28::N; i++) {
        unsigned long long s = (unsigned long long)Fq28C::p(i) * K + carry;
        kp[i] = (i < Fq28::N - 1) ? ((uint32_t)s & Fq28::MASK) : (uint32_t)s;
        carry = s >> 28;
    }
    // d = a − K·p with a borrow chain; keep it only if it did not go negative
    uint32_t d[Fq28::N];
    uint32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < Fq28::N - 1; i++) {
        uint32_t t = a.v[i] - kp[i] - borrow;
        borrow = t >> 31;                           // limbs < 2^28: a negative difference has bit 31 set
        d[i] = t & Fq28::MASK;
    }
    const int32_t top = (int32_t)(a.v[Fq28::N - 1] - kp[Fq28::N - 1] - borrow);
    if (top >= 0) {
#pragma unroll
        for (int i = 0; i < Fq28::N - 1; i++) a.v[i] = d[i];
        a.v[Fq28::N - 1] = (uint32_t)top;
    }
}
// fully canonical (< p, normalised limbs) from any value < 32p
FF_DEV Fq28 fq28_canon(const Fq28& x) {
    Fq28 a = fq28_norm(x);
    fq28_cond_sub_kp(a, 16); fq28_cond_sub_kp(a, 8); fq28_cond_sub_kp(a, 4); fq28_cond_sub_kp(a, 2); fq28_cond_sub_kp(a, 1);
    return a;
}
// value ≡ 0 mod p ?  for value < 32p.  Fast reject on the low limb: k·p ≡ k mod 2^28.
FF_DEV bool fq28_is_zero_mod_p(const Fq28& x) {
    if ((x.v[0] & Fq28::MASK) >= 32u) return false;
    Fq28 a = fq28_canon(x);
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < Fq28::N; i++) t |= a.v[i];
    return t == 0;
}

// ---- repacking between 12 × 32-bit and 14 × 28-bit limbs (same integer) ----
FF_DEV Fq28 fq28_repack_from32(const Fq& a) {
    Fq28 r;
#pragma unroll
    for (int k = 0; k < Fq28::N; k++) {
        const int bit = 28 * k, w = bit >> 5, sh = bit & 31;
        uint32_t lo = a.v[w];
        uint32_t hi = (w + 1 < 12) ? a.v[w + 1] : 0u;
        r.v[k] = __funnelshift_r(lo, hi, sh) & Fq28::MASK;
    }
    return r;
}
FF_DEV Fq fq28_repack_to32(const Fq28& a) {     // a normalised, value < 2^384
    Fq r;
#pragma unroll
    for (int w = 0; w < 12; w++) {
        const int bit = 32 * w, k = bit / 28, o = bit - 28 * k;
        uint32_t x = a.v[k] >> o;
        if (k + 1 < Fq28::N) x |= a.v[k + 1] << (28 - o);
        if (k + 2 < Fq28::N && 56 - o < 32) x |= a.v[k + 2] << (56 - o);
        r.v[w] = x;
    }
    return r;
}
// reference Montgomery form (canonical, 12 × u32) → internal form, canonical
FF_DEV Fq28 fq28_from_ref(const Fq& a) {
    Fq28 c;
#pragma unroll
    for (int i = 0; i < Fq28::N; i++) c.v[i] = Fq28C::c_in(i);
    Fq28 r = fq28_mul(fq28_repack_from32(a), c);
    fq28_cond_sub_kp(r, 1);
    return r;
}
// internal form (limbs < 2^30, value < 32p) → reference Montgomery form, canonical
FF_DEV Fq fq28_to_ref(const Fq28& a) {
    Fq28 c;
#pragma unroll
    for (int i = 0; i < Fq28::N; i++) c.v[i] = Fq28C::c_out(i);
    Fq28 r = fq28_mul(a, c);
    fq28_cond_sub_kp(r, 1);
    return fq28_repack_to32(r);
}

}  // namespace b200

// =================================================================================================
// XYZZ accumulator over Fq28 (x = X/ZZ, y = Y/ZZZ).  Stored invariants between additions:
//   X, Y normalised limbs (< 2^28), value(X) < 10p, value(Y) < 4p; ZZ, ZZZ multiplication results.
// =================================================================================================
#include "../../snarkvm_b200/csrc/ec.cuh"

namespace b200 {

struct Affine28 {
    Fq28 x, y;      // internal form, canonical
    bool inf;
};
static constexpr int AFFINE28_WORDS = 28;     // 112 B, 16-byte aligned; bit 31 of word 13 = infinity flag

FF_DEV void store_affine28(uint32_t* p, const Affine28& a) {
    uint32_t w[AFFINE28_WORDS];
#pragma unroll
    for (int i = 0; i < 14; i++) { w[i] = a.x.v[i]; w[14 + i] = a.y.v[i]; }
    if (a.inf) w[13] |= 0x80000000u;
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < 7; i++) q[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}
FF_DEV Affine28 load_affine28(const uint32_t* p) {
    uint32_t w[AFFINE28_WORDS];
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < 7; i++) { uint4 t = __ldg(q + i); w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w; }
    Affine28 a;
    a.inf = (w[13] >> 31) != 0u;
    w[13] &= 0x7fffffffu;
#pragma unroll
    for (int i = 0; i < 14; i++) { a.x.v[i] = w[i]; a.y.v[i] = w[14 + i]; }
    return a;
}

struct XYZZ28 {
    Fq28 X, Y, ZZ, ZZZ;
    bool inf;

    FF_DEV static XYZZ28 infinity() {
        XYZZ28 r;
#pragma unroll
        for (int i = 0; i < 14; i++) { r.X.v[i] = 0; r.Y.v[i] = 0; r.ZZ.v[i] = 0; r.ZZZ.v[i] = 0; }
        r.inf = true;
        return r;
    }

    // slow path: this == ±q in affine terms
    FF_DEV void double_affine(const Fq28& qx, const Fq28& qy_canon) {
        // mdbl-2008-s-1 on canonical inputs, every intermediate brought back to canonical form
        Fq28 U = fq28_canon(fq28_add(qy_canon, qy_canon));
        Fq28 V = fq28_sqr(U);
        Fq28 W = fq28_mul(U, V);
        Fq28 S = fq28_mul(qx, V);
        Fq28 XX = fq28_sqr(qx);
        Fq28 M = fq28_canon(fq28_add(fq28_add(XX, XX), XX));
        Fq28 S2 = fq28_canon(fq28_add(S, S));
        Fq28 X3 = fq28_canon(fq28_sub_2p(fq28_sqr(M), S2));
        Fq28 t = fq28_canon(fq28_sub_2p(fq28_canon(S), X3));
        Fq28 Y3 = fq28_canon(fq28_sub_2p(fq28_mul(M, t), fq28_canon(fq28_mul(W, qy_canon))));
        X = X3; Y = Y3; ZZ = V; ZZZ = W; inf = false;
    }

    // this += (negate ? −q : q)   (madd-2008-s; bounds model-checked in tests/manual/fq28_model.py)
    FF_DEV void add_affine(const Affine28& q, bool negate) {
        if (q.inf) return;
        Fq28 qy = negate ? fq28_neg(q.y) : q.y;                 // limbs < 2^29, value ≤ 2p
        if (inf) { X = q.x; Y = negate ? fq28_canon(qy) : q.y; ZZ = fq28_const_one(); ZZZ = ZZ; inf = false; return; }
        Fq28 U2 = fq28_mul(q.x, ZZ);
        Fq28 S2 = fq28_mul(qy, ZZZ);
        Fq28 P = fq28_sub_16p(U2, X);                           // limbs < 2^29.6, value < 18p
        Fq28 R = fq28_sub_16p(S2, Y);
        if (fq28_is_zero_mod_p(P)) {
            if (fq28_is_zero_mod_p(R)) double_affine(q.x, negate ? fq28_canon(qy) : q.y);
            else *this = infinity();
            return;
        }
        Fq28 PP = fq28_sqr(P);
        Fq28 PPP = fq28_mul(P, PP);
        Fq28 Q = fq28_mul(X, PP);
        Fq28 X3 = fq28_norm(fq28_sub_8p_wide(fq28_sqr(R), fq28_add(fq28_add(PPP, Q), Q)));   // value < 10p
        Fq28 Y3 = fq28_norm(fq28_sub_2p(fq28_mul(R, fq28_sub_16p(Q, X3)), fq28_mul(Y, PPP)));  // value < 4p
        ZZ = fq28_mul(ZZ, PP);
        ZZZ = fq28_mul(ZZZ, PPP);
        X = X3; Y = Y3;
    }

    // → the 192-byte XYZZ image in the reference Montgomery form (what the reduction kernels and the host read)
    FF_DEV XYZZ to_ref() const {
        if (inf) return XYZZ::infinity();
        XYZZ r;
        r.X = fq28_to_ref(X); r.Y = fq28_to_ref(Y); r.ZZ = fq28_to_ref(ZZ); r.ZZZ = fq28_to_ref(ZZZ);
        return r;
    }
};

}  // namespace b200

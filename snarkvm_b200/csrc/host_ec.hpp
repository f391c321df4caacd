// Host-side BLS12-377 Fq / G1 arithmetic used ONLY for the O(windows) tail of an MSM:
// summing the per-window bucket sums with Horner doublings and normalising the
// result (one inversion).  This mirrors where the reference's own CUDA plugin
// finishes on the host (algorithms/cuda/cuda/snarkvm.cu:290-295 adds the per-GPU
// partial points with point_t::dadd on the CPU).  ~253 doublings + ≤ 24 additions;
// all O(n) work stays on the GPU.  Independent of oracle/ (which is test-only).
#pragma once
#include <cstdint>
#include <cstring>

namespace b200 { namespace host {

typedef unsigned __int128 u128;

struct Fq {
    uint64_t l[6];
};
static const uint64_t Q_MOD[6] = {0x8508c00000000001ull, 0x170b5d4430000000ull, 0x1ef3622fba094800ull,
                                  0x1a22d9f300f5138full, 0xc63b05c06ca1493bull, 0x01ae3a4617c510eaull};
static const uint64_t Q_R1[6] = {0x02cdffffffffff68ull, 0x51409f837fffffb1ull, 0x9f7db3a98a7d3ff2ull,
                                 0x7b4e97b76e7c6305ull, 0x4cf495bf803c84e8ull, 0x008d6661e2fdf49aull};
static const uint64_t Q_INV = 9586122913090633727ull;   // fq.rs:111

inline Fq fq_zero() { Fq r; memset(r.l, 0, 48); return r; }
inline Fq fq_one() { Fq r; memcpy(r.l, Q_R1, 48); return r; }
inline bool fq_is_zero(const Fq& a) { uint64_t t = 0; for (int i = 0; i < 6; i++) t |= a.l[i]; return t == 0; }
inline bool fq_eq(const Fq& a, const Fq& b) { return memcmp(a.l, b.l, 48) == 0; }
inline bool ge_mod(const uint64_t* t) {
    for (int i = 5; i >= 0; i--) { if (t[i] != Q_MOD[i]) return t[i] > Q_MOD[i]; }
    return true;
}
inline void sub_mod(uint64_t* t) {
    uint64_t br = 0;
    for (int i = 0; i < 6; i++) { u128 d = (u128)t[i] - Q_MOD[i] - br; t[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
}
inline Fq fq_add(const Fq& a, const Fq& b) {
    Fq r; u128 c = 0;
    for (int i = 0; i < 6; i++) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    if (ge_mod(r.l)) sub_mod(r.l);
    return r;
}
inline Fq fq_sub(const Fq& a, const Fq& b) {
    Fq r; uint64_t br = 0;
    for (int i = 0; i < 6; i++) { u128 d = (u128)a.l[i] - b.l[i] - br; r.l[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
    if (br) { u128 c = 0; for (int i = 0; i < 6; i++) { c += (u128)r.l[i] + Q_MOD[i]; r.l[i] = (uint64_t)c; c >>= 64; } }
    return r;
}
inline Fq fq_dbl(const Fq& a) { return fq_add(a, a); }
inline Fq fq_mul(const Fq& a, const Fq& b) {
    uint64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 6; i++) {
        u128 c = 0;
        for (int j = 0; j < 6; j++) { c += (u128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[6]; t[6] = (uint64_t)c; t[7] = (uint64_t)(c >> 64);
        uint64_t k = t[0] * Q_INV;
        c = (u128)k * Q_MOD[0] + t[0]; c >>= 64;
        for (int j = 1; j < 6; j++) { c += (u128)k * Q_MOD[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[6]; t[5] = (uint64_t)c; t[6] = t[7] + (uint64_t)(c >> 64);
    }
    if (t[6] || ge_mod(t)) sub_mod(t);
    Fq r; memcpy(r.l, t, 48); return r;
}
inline Fq fq_sqr(const Fq& a) { return fq_mul(a, a); }
inline Fq fq_inverse(const Fq& a) {            // a^{q-2}
    uint64_t e[6]; memcpy(e, Q_MOD, 48); e[0] -= 2;
    Fq acc = fq_one(); bool started = false;
    for (int i = 5; i >= 0; i--)
        for (int b = 63; b >= 0; b--) {
            if (started) acc = fq_sqr(acc);
            if ((e[i] >> b) & 1) { acc = started ? fq_mul(acc, a) : a; started = true; }
        }
    return acc;
}

struct Xyzz {
    Fq X, Y, ZZ, ZZZ;
};
inline Xyzz xyzz_inf() { Xyzz r; r.X = r.Y = r.ZZ = r.ZZZ = fq_zero(); return r; }
inline bool xyzz_is_inf(const Xyzz& p) { return fq_is_zero(p.ZZ); }
inline void xyzz_dbl(Xyzz& p) {
    if (xyzz_is_inf(p)) return;
    Fq U = fq_dbl(p.Y), V = fq_sqr(U), W = fq_mul(U, V), S = fq_mul(p.X, V);
    Fq XX = fq_sqr(p.X), M = fq_add(fq_dbl(XX), XX);
    Fq X3 = fq_sub(fq_sqr(M), fq_dbl(S));
    Fq Y3 = fq_sub(fq_mul(M, fq_sub(S, X3)), fq_mul(W, p.Y));
    p.X = X3; p.Y = Y3; p.ZZ = fq_mul(V, p.ZZ); p.ZZZ = fq_mul(W, p.ZZZ);
}
inline void xyzz_add(Xyzz& p, const Xyzz& o) {
    if (xyzz_is_inf(o)) return;
    if (xyzz_is_inf(p)) { p = o; return; }
    Fq U1 = fq_mul(p.X, o.ZZ), U2 = fq_mul(o.X, p.ZZ), S1 = fq_mul(p.Y, o.ZZZ), S2 = fq_mul(o.Y, p.ZZZ);
    Fq P = fq_sub(U2, U1), R = fq_sub(S2, S1);
    if (fq_is_zero(P)) { if (fq_is_zero(R)) xyzz_dbl(p); else p = xyzz_inf(); return; }
    Fq PP = fq_sqr(P), PPP = fq_mul(P, PP), Q = fq_mul(U1, PP);
    Fq X3 = fq_sub(fq_sub(fq_sqr(R), PPP), fq_dbl(Q));
    p.Y = fq_sub(fq_mul(R, fq_sub(Q, X3)), fq_mul(S1, PPP));
    p.X = X3;
    p.ZZ = fq_mul(fq_mul(p.ZZ, o.ZZ), PP);
    p.ZZZ = fq_mul(fq_mul(p.ZZZ, o.ZZZ), PPP);
}
// Writes the 144-byte image of `result.to_affine().to_projective()`:
// (x, y, R) or (0, R, 0) for infinity — projective.rs:51-54, 507-512; affine.rs:331-353.
inline void xyzz_to_normalised_projective(const Xyzz& p, uint64_t out[18]) {
    Fq one = fq_one();
    if (xyzz_is_inf(p)) { memset(out, 0, 144); memcpy(out + 6, one.l, 48); return; }
    Fq i = fq_inverse(fq_mul(p.ZZ, p.ZZZ));
    Fq x = fq_mul(p.X, fq_mul(i, p.ZZZ)), y = fq_mul(p.Y, fq_mul(i, p.ZZ));
    memcpy(out, x.l, 48); memcpy(out + 6, y.l, 48); memcpy(out + 12, one.l, 48);
}
// Jacobian (X, Y, Z) image -> XYZZ (ZZ = Z^2, ZZZ = Z^3)
inline Xyzz xyzz_from_projective(const uint64_t in[18]) {
    Xyzz p; memcpy(p.X.l, in, 48); memcpy(p.Y.l, in + 6, 48);
    Fq Z; memcpy(Z.l, in + 12, 48);
    p.ZZ = fq_sqr(Z); p.ZZZ = fq_mul(p.ZZ, Z);
    if (fq_is_zero(Z)) p = xyzz_inf();
    return p;
}
// Σ_w 2^{c·w} · window_sum[w]  (Horner from the top window; batched.rs:404-413)
inline Xyzz horner_windows(const Xyzz* sums, int nwin, int c) {
    Xyzz total = xyzz_inf();
    for (int w = nwin - 1; w >= 0; w--) {
        for (int k = 0; k < c; k++) xyzz_dbl(total);
        xyzz_add(total, sums[w]);
    }
    return total;
}

}}  // namespace b200::host

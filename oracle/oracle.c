/*
 * TEST INFRASTRUCTURE — CPU restatement ("port") of snarkVM's hot path.
 *
 * This file is the parity oracle and the timed CPU baseline.  It is NOT part
 * of the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load the library built from it.
 * Nothing under snarkvm_b200/ links or calls it.
 *
 * Parity status: PINNED.  Checked by tests/test_oracle_golden.py against the
 * reference's own constants and known-answer data (Fr/Fq Montgomery constants,
 * POWERS_OF_ROOTS_OF_UNITY table, circuit_0 domain elements, the G1 generator,
 * real points of powers-of-beta-15.usrs) and against the independent Python
 * big-int restatement oracle/bls12_377.py.  The Rust reference itself cannot
 * be compiled here (no cargo/rustc), so there is no oracle/_ref.
 *
 * Each function cites the reference file:line (relative to /root/reference)
 * whose algorithm it restates.  OpenMP stands in for rayon.
 *
 * Build: see oracle/Makefile  (gcc -O3 -march=native -fopenmp -shared -fPIC).
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
#define INL static inline __attribute__((always_inline))

/* ------------------------------------------------------------------ */
/* Limb helpers — utilities/src/biginteger/mod.rs:99-147              */
/* ------------------------------------------------------------------ */
INL int bn_is_zero(const uint64_t *a, int n) { uint64_t t = 0; for (int i = 0; i < n; i++) t |= a[i]; return t == 0; }
INL int bn_eq(const uint64_t *a, const uint64_t *b, int n) { uint64_t t = 0; for (int i = 0; i < n; i++) t |= a[i] ^ b[i]; return t == 0; }
INL int bn_cmp(const uint64_t *a, const uint64_t *b, int n) {
    for (int i = n - 1; i >= 0; i--) { if (a[i] < b[i]) return -1; if (a[i] > b[i]) return 1; }
    return 0;
}
INL uint64_t bn_add(uint64_t *r, const uint64_t *a, const uint64_t *b, int n) {
    u128 c = 0; for (int i = 0; i < n; i++) { c += (u128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; } return (uint64_t)c;
}
INL uint64_t bn_sub(uint64_t *r, const uint64_t *a, const uint64_t *b, int n) {
    uint64_t br = 0;
    for (int i = 0; i < n; i++) { u128 t = (u128)a[i] - b[i] - br; r[i] = (uint64_t)t; br = (uint64_t)(t >> 64) & 1; }
    return br;
}
INL void bn_div2(uint64_t *a, int n) { for (int i = 0; i < n - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 63); a[n - 1] >>= 1; }
INL int bn_is_one(const uint64_t *a, int n) { if (a[0] != 1) return 0; for (int i = 1; i < n; i++) if (a[i]) return 0; return 1; }

/* ------------------------------------------------------------------ */
/* Generic Montgomery field on n 64-bit limbs.                        */
/* fields/src/fp_256.rs:52-65,730-817 ; fields/src/fp_384.rs:52-65,   */
/* 771-899.  Values are kept fully reduced (reduce(), :61-65).        */
/* ------------------------------------------------------------------ */
INL void fp_add(uint64_t *r, const uint64_t *a, const uint64_t *b, const uint64_t *m, int n) {
    bn_add(r, a, b, n);                       /* modulus has spare bits: no carry out */
    if (bn_cmp(r, m, n) >= 0) bn_sub(r, r, m, n);
}
INL void fp_sub(uint64_t *r, const uint64_t *a, const uint64_t *b, const uint64_t *m, int n) {
    uint64_t t[6];
    if (bn_cmp(b, a, n) > 0) { bn_add(t, a, m, n); bn_sub(r, t, b, n); } else bn_sub(r, a, b, n);
}
INL void fp_neg(uint64_t *r, const uint64_t *a, const uint64_t *m, int n) {
    if (bn_is_zero(a, n)) { for (int i = 0; i < n; i++) r[i] = 0; } else bn_sub(r, m, a, n);
}
/* CIOS multiplication — fp_256.rs:754-817, fp_384.rs:771-899 */
INL void fp_mul(uint64_t *r, const uint64_t *a, const uint64_t *b, const uint64_t *m, uint64_t inv, int n) {
    uint64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; i++) {
        u128 c = 0;
        for (int j = 0; j < n; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[n]; t[n] = (uint64_t)c; t[n + 1] = (uint64_t)(c >> 64);
        uint64_t k = t[0] * inv;
        c = (u128)k * m[0] + t[0]; c >>= 64;
        for (int j = 1; j < n; j++) { c += (u128)k * m[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[n]; t[n - 1] = (uint64_t)c; t[n] = t[n + 1] + (uint64_t)(c >> 64);
    }
    if (t[n] || bn_cmp(t, m, n) >= 0) bn_sub(t, t, m, n);
    for (int i = 0; i < n; i++) r[i] = t[i];
}
/* Binary extended Euclid ("BEA for inversion in Fp") — fp_256.rs:290-340, fp_384.rs:348-398.
 * Input and output in Montgomery form (b starts at R2).  Returns 0 if a == 0. */
INL int fp_inverse(uint64_t *r, const uint64_t *a, const uint64_t *m, const uint64_t *r2, int n) {
    if (bn_is_zero(a, n)) return 0;
    uint64_t u[6], v[6], b[6], c[6];
    for (int i = 0; i < n; i++) { u[i] = a[i]; v[i] = m[i]; b[i] = r2[i]; c[i] = 0; }
    while (!bn_is_one(u, n) && !bn_is_one(v, n)) {
        while (!(u[0] & 1)) {
            bn_div2(u, n);
            if (!(b[0] & 1)) bn_div2(b, n);
            else { uint64_t cy = bn_add(b, b, m, n); bn_div2(b, n); b[n - 1] |= cy << 63; }
        }
        while (!(v[0] & 1)) {
            bn_div2(v, n);
            if (!(c[0] & 1)) bn_div2(c, n);
            else { uint64_t cy = bn_add(c, c, m, n); bn_div2(c, n); c[n - 1] |= cy << 63; }
        }
        if (bn_cmp(v, u, n) < 0) { bn_sub(u, u, v, n); fp_sub(b, b, c, m, n); }
        else { bn_sub(v, v, u, n); fp_sub(c, c, b, m, n); }
    }
    const uint64_t *s = bn_is_one(u, n) ? b : c;
    for (int i = 0; i < n; i++) r[i] = s[i];
    return 1;
}

/* ------------------------------------------------------------------ */
/* Fr — curves/src/bls12_377/fr.rs:109-192                            */
/* ------------------------------------------------------------------ */
typedef struct { uint64_t l[4]; } fr_t;
static const uint64_t FR_MOD[4] = {725501752471715841ull, 6461107452199829505ull, 6968279316240510977ull, 1345280370688173398ull};
static const uint64_t FR_R[4]   = {9015221291577245683ull, 8239323489949974514ull, 1646089257421115374ull, 958099254763297437ull};
static const uint64_t FR_R2[4]  = {2726216793283724667ull, 14712177743343147295ull, 12091039717619697043ull, 81024008013859129ull};
static const uint64_t FR_INV    = 725501752471715839ull;
/* TWO_ADIC_ROOT_OF_UNITY (Montgomery), fr.rs:115-120; TWO_ADICITY = 47 (:109) */
static const uint64_t FR_ROOT47[4] = {12646347781564978760ull, 6783048705277173164ull, 268534165941069093ull, 1121515446318641358ull};
/* GENERATOR = 22 (Montgomery), fr.rs:130-135 */
static const uint64_t FR_GEN[4] = {2984901390528151251ull, 10561528701063790279ull, 5476750214495080041ull, 898978044469942640ull};

INL void fr_add(fr_t *r, const fr_t *a, const fr_t *b) { fp_add(r->l, a->l, b->l, FR_MOD, 4); }
INL void fr_sub(fr_t *r, const fr_t *a, const fr_t *b) { fp_sub(r->l, a->l, b->l, FR_MOD, 4); }
INL void fr_mul(fr_t *r, const fr_t *a, const fr_t *b) { fp_mul(r->l, a->l, b->l, FR_MOD, FR_INV, 4); }
INL void fr_sqr(fr_t *r, const fr_t *a) { fp_mul(r->l, a->l, a->l, FR_MOD, FR_INV, 4); }
static int fr_inverse(fr_t *r, const fr_t *a) { return fp_inverse(r->l, a->l, FR_MOD, FR_R2, 4); }
INL int fr_is_zero(const fr_t *a) { return bn_is_zero(a->l, 4); }
static fr_t fr_one(void) { fr_t o; memcpy(o.l, FR_R, 32); return o; }
static fr_t fr_pow_u64(const fr_t *a, uint64_t e) {
    fr_t acc = fr_one(), b = *a;
    while (e) { if (e & 1) fr_mul(&acc, &acc, &b); fr_sqr(&b, &b); e >>= 1; }
    return acc;
}
static fr_t fr_from_u64(uint64_t v) { fr_t t = {{v, 0, 0, 0}}, r2; memcpy(r2.l, FR_R2, 32); fr_mul(&t, &t, &r2); return t; }

/* ------------------------------------------------------------------ */
/* Fq — curves/src/bls12_377/fq.rs:85-176                             */
/* ------------------------------------------------------------------ */
typedef struct { uint64_t l[6]; } fq_t;
static const uint64_t FQ_MOD[6] = {0x8508c00000000001ull, 0x170b5d4430000000ull, 0x1ef3622fba094800ull, 0x1a22d9f300f5138full, 0xc63b05c06ca1493bull, 0x1ae3a4617c510eaull};
static const uint64_t FQ_R[6]   = {202099033278250856ull, 5854854902718660529ull, 11492539364873682930ull, 8885205928937022213ull, 5545221690922665192ull, 39800542322357402ull};
static const uint64_t FQ_R2[6]  = {0xb786686c9400cd22ull, 0x329fcaab00431b1ull, 0x22a5f11162d6b46dull, 0xbfdf7d03827dc3acull, 0x837e92f041790bf9ull, 0x6dfccb1e914b88ull};
static const uint64_t FQ_INV    = 9586122913090633727ull;

INL void fq_add(fq_t *r, const fq_t *a, const fq_t *b) { fp_add(r->l, a->l, b->l, FQ_MOD, 6); }
INL void fq_sub(fq_t *r, const fq_t *a, const fq_t *b) { fp_sub(r->l, a->l, b->l, FQ_MOD, 6); }
INL void fq_neg(fq_t *r, const fq_t *a) { fp_neg(r->l, a->l, FQ_MOD, 6); }
INL void fq_dbl(fq_t *r, const fq_t *a) { fp_add(r->l, a->l, a->l, FQ_MOD, 6); }
INL void fq_mul(fq_t *r, const fq_t *a, const fq_t *b) { fp_mul(r->l, a->l, b->l, FQ_MOD, FQ_INV, 6); }
INL void fq_sqr(fq_t *r, const fq_t *a) { fp_mul(r->l, a->l, a->l, FQ_MOD, FQ_INV, 6); }
static int fq_inverse(fq_t *r, const fq_t *a) { return fp_inverse(r->l, a->l, FQ_MOD, FQ_R2, 6); }
INL int fq_is_zero(const fq_t *a) { return bn_is_zero(a->l, 6); }
INL int fq_eq(const fq_t *a, const fq_t *b) { return bn_eq(a->l, b->l, 6); }
static fq_t fq_one(void) { fq_t o; memcpy(o.l, FQ_R, 48); return o; }
static fq_t fq_zero(void) { fq_t o; memset(o.l, 0, 48); return o; }
/* half() = (q+1)/2 in Montgomery form — fp_384.rs:191-198 */
static fq_t fq_half(void) {
    fq_t t, r2; memcpy(t.l, FQ_MOD, 48); t.l[0] += 1; bn_div2(t.l, 6);
    memcpy(r2.l, FQ_R2, 48); fq_mul(&t, &t, &r2); return t;
}

/* ------------------------------------------------------------------ */
/* G1 — curves/src/templates/short_weierstrass_jacobian/{affine,      */
/* projective}.rs ; curve y^2 = x^3 + 1 (bls12_377/g1.rs:78-91)       */
/* ------------------------------------------------------------------ */
typedef struct { fq_t x, y; uint8_t inf; uint8_t pad[7]; } g1_affine_t;   /* 104 B, affine.rs:41-46 */
typedef struct { fq_t x, y, z; } g1_proj_t;                                 /* 144 B, projective.rs:36-41 */

static g1_affine_t aff_zero(void) { g1_affine_t a; memset(&a, 0, sizeof a); a.y = fq_one(); a.inf = 1; return a; } /* affine.rs:57-59 */
static g1_proj_t proj_zero(void) { g1_proj_t p; p.x = fq_zero(); p.y = fq_one(); p.z = fq_zero(); return p; }        /* projective.rs:51-54 */
INL int proj_is_zero(const g1_proj_t *p) { return fq_is_zero(&p->z); }

/* double_in_place, a = 0 branch — projective.rs:302-339 */
static void proj_double(g1_proj_t *p) {
    if (proj_is_zero(p)) return;
    fq_t a, b, c, d, e, f, t;
    fq_sqr(&a, &p->x); fq_sqr(&b, &p->y); fq_sqr(&c, &b);
    fq_add(&t, &p->x, &b); fq_sqr(&t, &t); fq_sub(&t, &t, &a); fq_sub(&t, &t, &c); fq_dbl(&d, &t);
    fq_dbl(&e, &a); fq_add(&e, &e, &a);
    fq_sqr(&f, &e);
    fq_mul(&p->z, &p->z, &p->y); fq_dbl(&p->z, &p->z);
    fq_dbl(&t, &d); fq_sub(&p->x, &f, &t);
    fq_dbl(&c, &c); fq_dbl(&c, &c); fq_dbl(&c, &c);
    fq_sub(&t, &d, &p->x); fq_mul(&t, &t, &e); fq_sub(&p->y, &t, &c);
}
/* add_assign_mixed (madd-2007-bl) — projective.rs:222-291 */
static void proj_add_mixed(g1_proj_t *p, const g1_affine_t *q) {
    if (q->inf) return;
    if (proj_is_zero(p)) { p->x = q->x; p->y = q->y; p->z = fq_one(); return; }
    fq_t z1z1, u2, s2, h, hh, i, j, r, v, t, t2;
    fq_sqr(&z1z1, &p->z);
    fq_mul(&u2, &q->x, &z1z1);
    fq_mul(&s2, &q->y, &p->z); fq_mul(&s2, &s2, &z1z1);
    if (fq_eq(&p->x, &u2) && fq_eq(&p->y, &s2)) { proj_double(p); return; }
    fq_sub(&h, &u2, &p->x);
    fq_sqr(&hh, &h);
    fq_dbl(&i, &hh); fq_dbl(&i, &i);
    fq_mul(&j, &h, &i);
    fq_sub(&r, &s2, &p->y); fq_dbl(&r, &r);
    fq_mul(&v, &p->x, &i);
    fq_t x3; fq_sqr(&x3, &r); fq_sub(&x3, &x3, &j); fq_dbl(&t, &v); fq_sub(&x3, &x3, &t);
    /* Y3 = r*(V-X3) - 2*Y1*J   (sum_of_products in the reference, fp_384.rs:200-262) */
    fq_sub(&t, &v, &x3); fq_mul(&t, &r, &t);
    fq_dbl(&t2, &p->y); fq_mul(&t2, &t2, &j);
    fq_sub(&p->y, &t, &t2);
    p->x = x3;
    fq_add(&t, &p->z, &h); fq_sqr(&t, &t); fq_sub(&t, &t, &z1z1); fq_sub(&p->z, &t, &hh);
}
/* add_assign (add-2007-bl) — projective.rs:407-468 */
static void proj_add(g1_proj_t *p, const g1_proj_t *q) {
    if (proj_is_zero(p)) { *p = *q; return; }
    if (proj_is_zero(q)) return;
    fq_t z1z1, z2z2, u1, u2, s1, s2, h, i, j, r, v, t, t2;
    fq_sqr(&z1z1, &p->z); fq_sqr(&z2z2, &q->z);
    fq_mul(&u1, &p->x, &z2z2); fq_mul(&u2, &q->x, &z1z1);
    fq_mul(&s1, &p->y, &q->z); fq_mul(&s1, &s1, &z2z2);
    fq_mul(&s2, &q->y, &p->z); fq_mul(&s2, &s2, &z1z1);
    if (fq_eq(&u1, &u2) && fq_eq(&s1, &s2)) { proj_double(p); return; }
    fq_sub(&h, &u2, &u1);
    fq_dbl(&i, &h); fq_sqr(&i, &i);
    fq_mul(&j, &h, &i);
    fq_sub(&r, &s2, &s1); fq_dbl(&r, &r);
    fq_mul(&v, &u1, &i);
    fq_t x3; fq_sqr(&x3, &r); fq_sub(&x3, &x3, &j); fq_dbl(&t, &v); fq_sub(&x3, &x3, &t);
    fq_sub(&t, &v, &x3); fq_mul(&t, &r, &t);
    fq_dbl(&t2, &s1); fq_mul(&t2, &t2, &j);
    fq_sub(&p->y, &t, &t2);
    p->x = x3;
    fq_add(&t, &p->z, &q->z); fq_sqr(&t, &t); fq_sub(&t, &t, &z1z1); fq_sub(&t, &t, &z2z2); fq_mul(&p->z, &t, &h);
}
/* From<Projective> for Affine — affine.rs:331-353 */
static g1_affine_t proj_to_affine(const g1_proj_t *p) {
    if (proj_is_zero(p)) return aff_zero();
    g1_affine_t a; memset(&a, 0, sizeof a);
    fq_t one = fq_one();
    if (fq_eq(&p->z, &one)) { a.x = p->x; a.y = p->y; return a; }
    fq_t zi, zi2, zi3; fq_inverse(&zi, &p->z); fq_sqr(&zi2, &zi); fq_mul(&zi3, &zi2, &zi);
    fq_mul(&a.x, &p->x, &zi2); fq_mul(&a.y, &p->y, &zi3);
    return a;
}
/* From<Affine> for Projective — projective.rs:507-512 */
static g1_proj_t aff_to_proj(const g1_affine_t *a) {
    if (a->inf) return proj_zero();
    g1_proj_t p; p.x = a->x; p.y = a->y; p.z = fq_one(); return p;
}
/* mul_bits over BitIteratorBE of a canonical 4-limb scalar — affine.rs:173-182 */
static g1_proj_t aff_mul_bits(const g1_affine_t *a, const uint64_t s[4]) {
    g1_proj_t out = proj_zero();
    int started = 0;
    for (int bit = 255; bit >= 0; bit--) {
        int b = (s[bit >> 6] >> (bit & 63)) & 1;
        if (!started) { if (!b) continue; started = 1; }
        proj_double(&out);
        if (b) proj_add_mixed(&out, a);
    }
    return out;
}
/* batch_add_loop_1 — affine.rs:224-254 */
INL void batch_add_loop_1(g1_affine_t *a, g1_affine_t *b, const fq_t *half, fq_t *inv_tmp) {
    if (a->inf || b->inf) return;
    if (fq_eq(&a->x, &b->x)) {
        if (fq_eq(&a->y, &b->y)) {
            fq_t x_sq, t; fq_sqr(&x_sq, &b->x);
            fq_sub(&b->x, &b->x, &b->y);
            fq_dbl(&a->x, &b->y);
            fq_dbl(&t, &x_sq); fq_add(&a->y, &t, &x_sq);      /* + WEIERSTRASS_A (= 0) */
            fq_mul(&t, &a->y, half); fq_sub(&b->y, &b->y, &t);
            fq_mul(&a->y, &a->y, inv_tmp);
            fq_mul(inv_tmp, inv_tmp, &a->x);
        } else { a->inf = 1; b->inf = 1; }
    } else {
        fq_sub(&a->x, &a->x, &b->x);
        fq_sub(&a->y, &a->y, &b->y);
        fq_mul(&a->y, &a->y, inv_tmp);
        fq_mul(inv_tmp, inv_tmp, &a->x);
    }
}
/* batch_add_loop_2 — affine.rs:259-273 */
INL void batch_add_loop_2(g1_affine_t *a, const g1_affine_t *b, fq_t *inv_tmp) {
    if (a->inf) { *a = *b; return; }
    if (b->inf) return;
    fq_t lambda, t;
    fq_mul(&lambda, &a->y, inv_tmp);
    fq_mul(inv_tmp, inv_tmp, &a->x);
    fq_dbl(&t, &b->x); fq_add(&a->x, &a->x, &t);
    fq_sqr(&t, &lambda); fq_sub(&a->x, &t, &a->x);
    fq_sub(&t, &b->x, &a->x); fq_mul(&t, &lambda, &t); fq_sub(&a->y, &t, &b->y);
}

/* ------------------------------------------------------------------ */
/* batched::msm — algorithms/src/msm/variable_base/batched.rs          */
/* ------------------------------------------------------------------ */
typedef struct { uint32_t bucket, idx; } bucket_pos_t;          /* BucketPosition, :26-29 */
typedef struct { uint32_t a, b; } instr_t;

/* batch_size — batched.rs:56-73 (x86_64 branch) */
static size_t batch_size_for(size_t n) { return n < 500000 ? 300 : 3000; }

/* batch_add_write — batched.rs:131-172.  scratch holds the `b` halves. */
static void batch_add_write(const uint8_t *bases, size_t stride, const instr_t *ins, size_t nins,
                            g1_affine_t *out, size_t *out_len, g1_affine_t *scratch, uint8_t *has_b,
                            const fq_t *half) {
    fq_t inv_tmp = fq_one();
    size_t base_len = *out_len;
    for (size_t i = 0; i < nins; i++) {
        g1_affine_t a; memcpy(&a, bases + (size_t)ins[i].a * stride, 104); a.inf = a.inf != 0;
        if (ins[i].b == 0xFFFFFFFFu) { out[base_len + i] = a; has_b[i] = 0; }
        else {
            g1_affine_t b; memcpy(&b, bases + (size_t)ins[i].b * stride, 104); b.inf = b.inf != 0;
            batch_add_loop_1(&a, &b, half, &inv_tmp);
            out[base_len + i] = a; scratch[i] = b; has_b[i] = 1;
        }
    }
    fq_inverse(&inv_tmp, &inv_tmp);
    for (size_t i = nins; i-- > 0;)
        if (has_b[i]) batch_add_loop_2(&out[base_len + i], &scratch[i], &inv_tmp);
    *out_len = base_len + nins;
}
/* batch_add_in_place_same_slice — batched.rs:78-122 */
static void batch_add_in_place(g1_affine_t *bases, const instr_t *ins, size_t nins, const fq_t *half) {
    fq_t inv_tmp = fq_one();
    for (size_t i = 0; i < nins; i++) batch_add_loop_1(&bases[ins[i].a], &bases[ins[i].b], half, &inv_tmp);
    fq_inverse(&inv_tmp, &inv_tmp);
    for (size_t i = nins; i-- > 0;) { g1_affine_t b = bases[ins[i].b]; batch_add_loop_2(&bases[ins[i].a], &b, &inv_tmp); }
}
/* sort_unstable by bucket_index (batched.rs:187) — LSD radix sort, 2 × 16-bit digits
 * over the (≤ 2^c ≤ 2^22 here, but handle full 32 bits) key. */
static void sort_positions(bucket_pos_t *p, size_t n, bucket_pos_t *tmp) {
    for (int pass = 0; pass < 2; pass++) {
        int shift = pass * 16;
        int needed = 0;
        for (size_t i = 0; i < n && !needed; i++) if ((p[i].bucket >> shift) & 0xFFFF) needed = 1;
        if (!needed && pass == 1) break;
        size_t *cnt = (size_t *)calloc(65537, sizeof(size_t));
        for (size_t i = 0; i < n; i++) cnt[((p[i].bucket >> shift) & 0xFFFF) + 1]++;
        for (int k = 0; k < 65536; k++) cnt[k + 1] += cnt[k];
        for (size_t i = 0; i < n; i++) tmp[cnt[(p[i].bucket >> shift) & 0xFFFF]++] = p[i];
        memcpy(p, tmp, n * sizeof *p);
        free(cnt);
    }
}
/* batch_add — batched.rs:175-325.  Returns malloc'd res[num_buckets]. */
static g1_affine_t *batch_add(size_t num_buckets, const uint8_t *bases, size_t stride, size_t nbases,
                              bucket_pos_t *pos, size_t npos) {
    const fq_t half = fq_half();
    size_t bsz = batch_size_for(nbases);
    bucket_pos_t *tmp = (bucket_pos_t *)malloc((npos ? npos : 1) * sizeof *tmp);
    sort_positions(pos, npos, tmp);
    free(tmp);

    size_t num_scalars = npos, new_len = 0, gc = 0, lc = 1, in_batch = 0;
    int all_ones = 1;
    size_t icap = bsz + 8, nins = 0;
    instr_t *ins = (instr_t *)malloc(icap * sizeof *ins);
    g1_affine_t *new_bases = (g1_affine_t *)malloc((nbases ? nbases : 1) * sizeof *new_bases);
    size_t new_bases_len = 0;
    g1_affine_t *scratch = NULL; uint8_t *has_b = NULL; size_t scap = 0;
#define ENSURE_INS(extra) do { if (nins + (extra) > icap) { icap = (nins + (extra)) * 2; ins = (instr_t *)realloc(ins, icap * sizeof *ins); } } while (0)
#define FLUSH_WRITE() do { if (nins > scap) { scap = nins * 2; scratch = (g1_affine_t *)realloc(scratch, scap * sizeof *scratch); has_b = (uint8_t *)realloc(has_b, scap); } \
        batch_add_write(bases, stride, ins, nins, new_bases, &new_bases_len, scratch, has_b, &half); nins = 0; } while (0)

    while (gc < num_scalars) {
        uint32_t cur = pos[gc].bucket;
        while (gc + 1 < num_scalars && pos[gc + 1].bucket == cur) { gc++; lc++; }
        if (cur >= (uint32_t)num_buckets) { lc = 1; }
        else if (lc > 1) {
            if (lc > 2) all_ones = 0;
            size_t hf = lc / 2; int odd = lc & 1;
            ENSURE_INS(hf + 1);
            for (size_t i = 0; i < hf; i++) {
                ins[nins].a = pos[gc - (lc - 1) + 2 * i].idx; ins[nins].b = pos[gc - (lc - 1) + 2 * i + 1].idx; nins++;
                pos[new_len + i].bucket = cur; pos[new_len + i].idx = (uint32_t)(new_len + i);
            }
            if (odd) {
                ins[nins].a = pos[gc].idx; ins[nins].b = 0xFFFFFFFFu; nins++;
                pos[new_len + hf].bucket = cur; pos[new_len + hf].idx = (uint32_t)(new_len + hf);
            }
            new_len += hf + (lc & 1); in_batch += hf; lc = 1;
            if (in_batch >= bsz / 2) { FLUSH_WRITE(); in_batch = 0; }
        } else {
            ENSURE_INS(1);
            ins[nins].a = pos[gc].idx; ins[nins].b = 0xFFFFFFFFu; nins++;
            pos[new_len].bucket = cur; pos[new_len].idx = (uint32_t)new_len; new_len++;
        }
        gc++;
    }
    if (nins) FLUSH_WRITE();
    gc = 0; in_batch = 0; lc = 1; num_scalars = new_len; new_len = 0;

    while (!all_ones) {
        all_ones = 1;
        while (gc < num_scalars) {
            uint32_t cur = pos[gc].bucket;
            while (gc + 1 < num_scalars && pos[gc + 1].bucket == cur) { gc++; lc++; }
            if (cur >= (uint32_t)num_buckets) { lc = 1; }
            else if (lc > 1) {
                if (lc != 2) all_ones = 0;
                size_t hf = lc / 2; int odd = lc & 1;
                ENSURE_INS(hf);
                for (size_t i = 0; i < hf; i++) {
                    ins[nins].a = pos[gc - (lc - 1) + 2 * i].idx; ins[nins].b = pos[gc - (lc - 1) + 2 * i + 1].idx; nins++;
                    pos[new_len + i] = pos[gc - (lc - 1) + 2 * i];
                }
                if (odd) pos[new_len + hf] = pos[gc];
                new_len += hf + (lc & 1); in_batch += hf; lc = 1;
                if (in_batch >= bsz / 2) { batch_add_in_place(new_bases, ins, nins, &half); nins = 0; in_batch = 0; }
            } else { pos[new_len] = pos[gc]; new_len++; }
            gc++;
        }
        if (nins) { batch_add_in_place(new_bases, ins, nins, &half); nins = 0; }
        gc = 0; in_batch = 0; lc = 1; num_scalars = new_len; new_len = 0;
    }
    g1_affine_t *res = (g1_affine_t *)malloc((num_buckets ? num_buckets : 1) * sizeof *res);
    g1_affine_t z = aff_zero();
    for (size_t i = 0; i < num_buckets; i++) res[i] = z;
    for (size_t i = 0; i < num_scalars; i++) res[pos[i].bucket] = new_bases[pos[i].idx];
    free(ins); free(new_bases); free(scratch); free(has_b);
    return res;
#undef ENSURE_INS
#undef FLUSH_WRITE
}
/* (scalar >> w_start) mod 2^c — BigInteger::divn + as_ref()[0] % (1 << c), batched.rs:341-347 */
INL uint64_t scalar_window(const uint64_t s[4], unsigned w_start, unsigned c) {
    unsigned limb = w_start >> 6, sh = w_start & 63;
    if (limb >= 4) return 0;
    uint64_t v = s[limb] >> sh;
    if (sh && limb + 1 < 4) v |= s[limb + 1] << (64 - sh);
    return c >= 64 ? v : (v & ((1ull << c) - 1));
}
/* batched_window — batched.rs:328-364 */
static g1_proj_t batched_window(const uint8_t *bases, size_t stride, size_t nbases, const uint64_t *scalars, size_t n,
                                unsigned w_start, unsigned c) {
    size_t num_buckets = ((size_t)1 << c) - 1;
    bucket_pos_t *pos = (bucket_pos_t *)malloc((n ? n : 1) * sizeof *pos);
    for (size_t i = 0; i < n; i++) {
        uint32_t d = (uint32_t)scalar_window(scalars + 4 * i, w_start, c);
        pos[i].bucket = d - 1u;          /* digit 0 wraps to 0xFFFFFFFF ≥ num_buckets ⇒ skipped (:207-208,350) */
        pos[i].idx = (uint32_t)i;
    }
    g1_affine_t *buckets = batch_add(num_buckets, bases, stride, nbases, pos, n);
    g1_proj_t res = proj_zero(), running = proj_zero();
    for (size_t b = num_buckets; b-- > 0;) { proj_add_mixed(&running, &buckets[b]); proj_add(&res, &running); }
    free(buckets); free(pos);
    return res;
}
/* ln_without_floats — algorithms/src/msm/mod.rs:29-32 ; log2 = ceil(log2), fft/domain.rs:64-72 */
static unsigned ceil_log2(size_t x) { if (x <= 1) return 0; unsigned l = 0; size_t v = x - 1; while (v) { l++; v >>= 1; } return l; }
static unsigned ln_without_floats(size_t a) { return ceil_log2(a) * 69 / 100; }

/* batched::msm — batched.rs:366-415 */
static g1_proj_t msm_batched(const uint8_t *bases, size_t stride, size_t nbases, const uint64_t *scalars, size_t n) {
    const unsigned num_bits = 253;
    if (nbases < 15) {
        /* bit-serial path, :367-387 */
        g1_proj_t sum = proj_zero(); int seen = 0;
        size_t m = n < nbases ? n : nbases;
        for (int bit = (int)num_bits - 1; bit >= 0; bit--) {
            if (seen) proj_double(&sum);
            for (size_t i = 0; i < m; i++)
                if ((scalars[4 * i + (bit >> 6)] >> (bit & 63)) & 1) {
                    g1_affine_t a; memcpy(&a, bases + i * stride, 104); a.inf = a.inf != 0;
                    proj_add_mixed(&sum, &a); seen = 1;
                }
        }
        return sum;
    }
    unsigned c = n < 32 ? 1 : ln_without_floats(n) + 2;
    unsigned nwin = (num_bits + c - 1) / c;
    g1_proj_t *sums = (g1_proj_t *)malloc(nwin * sizeof *sums);
#pragma omp parallel for schedule(dynamic, 1)
    for (unsigned w = 0; w < nwin; w++) sums[w] = batched_window(bases, stride, nbases, scalars, n, w * c, c);
    g1_proj_t total = proj_zero();
    for (unsigned w = nwin; w-- > 1;) { proj_add(&total, &sums[w]); for (unsigned k = 0; k < c; k++) proj_double(&total); }
    proj_add(&total, &sums[0]);
    free(sums);
    return total;
}

/* standard::msm — algorithms/src/msm/variable_base/standard.rs:23-105 */
static g1_proj_t msm_standard(const uint8_t *bases, size_t stride, const uint64_t *scalars, size_t n) {
    const unsigned num_bits = 253;
    unsigned c = n < 32 ? 3 : ln_without_floats(n) + 2;
    unsigned nwin = (num_bits + c - 1) / c;
    g1_proj_t *sums = (g1_proj_t *)malloc(nwin * sizeof *sums);
#pragma omp parallel for schedule(dynamic, 1)
    for (unsigned w = 0; w < nwin; w++) {
        unsigned w_start = w * c;
        g1_proj_t res = proj_zero();
        size_t nb = ((size_t)1 << c) - 1;
        g1_proj_t *buckets = (g1_proj_t *)malloc(nb * sizeof *buckets);
        for (size_t b = 0; b < nb; b++) buckets[b] = proj_zero();
        for (size_t i = 0; i < n; i++) {
            const uint64_t *s = scalars + 4 * i;
            g1_affine_t a; memcpy(&a, bases + i * stride, 104); a.inf = a.inf != 0;
            int is_one = s[0] == 1 && !s[1] && !s[2] && !s[3];
            if (is_one) { if (w_start == 0) proj_add_mixed(&res, &a); continue; }   /* :52-57 */
            uint64_t d = scalar_window(s, w_start, c);
            if (d != 0) proj_add_mixed(&buckets[d - 1], &a);
        }
        /* to affine then running sum, :68-76 (batch_normalization is a representation change only) */
        g1_proj_t running = proj_zero();
        for (size_t b = nb; b-- > 0;) { proj_add(&running, &buckets[b]); proj_add(&res, &running); }
        free(buckets);
        sums[w] = res;
    }
    g1_proj_t total = proj_zero();
    for (unsigned w = nwin; w-- > 1;) { proj_add(&total, &sums[w]); for (unsigned k = 0; k < c; k++) proj_double(&total); }
    proj_add(&total, &sums[0]);
    free(sums);
    return total;
}
/* msm_naive — variable_base/mod.rs:52-57 (parallel variant :59-66) */
static g1_proj_t msm_naive(const uint8_t *bases, size_t stride, const uint64_t *scalars, size_t n) {
    g1_proj_t acc = proj_zero();
#pragma omp parallel
    {
        g1_proj_t local = proj_zero();
#pragma omp for schedule(static)
        for (size_t i = 0; i < n; i++) {
            g1_affine_t a; memcpy(&a, bases + i * stride, 104); a.inf = a.inf != 0;
            g1_proj_t t = aff_mul_bits(&a, scalars + 4 * i);
            proj_add(&local, &t);
        }
#pragma omp critical
        proj_add(&acc, &local);
    }
    return acc;
}

/* ------------------------------------------------------------------ */
/* EvaluationDomain — algorithms/src/fft/domain.rs                     */
/* ------------------------------------------------------------------ */
/* get_root_of_unity — fields/src/traits/fft_field.rs:38-86: 2-adic root squared (47 - lg) times */
static fr_t fr_root_of_unity(unsigned lg) {
    fr_t w; memcpy(w.l, FR_ROOT47, 32);
    for (unsigned i = lg; i < 47; i++) fr_sqr(&w, &w);
    return w;
}
/* roots_of_unity(root): [1, g, …, g^{n/2-1}] — domain.rs:594-648 (chunked pow instead of the recursion) */
static fr_t *roots_of_unity(const fr_t *root, size_t half) {
    fr_t *r = (fr_t *)malloc((half ? half : 1) * sizeof *r);
    size_t chunk = 1024;
#pragma omp parallel for schedule(static)
    for (size_t s = 0; s < half; s += chunk) {
        fr_t p = fr_pow_u64(root, s);
        size_t e = s + chunk < half ? s + chunk : half;
        for (size_t i = s; i < e; i++) { r[i] = p; fr_mul(&p, &p, root); }
    }
    return r;
}
/* derange_helper / bitrev — domain.rs:789-804 */
static void derange(fr_t *x, unsigned lg) {
    size_t n = (size_t)1 << lg;
    if (n <= 2) return;
#pragma omp parallel for schedule(static)
    for (size_t i = 1; i < n - 1; i++) {
        size_t r = 0, v = i;
        for (unsigned b = 0; b < lg; b++) { r = (r << 1) | (v & 1); v >>= 1; }
        if (i < r) { fr_t t = x[i]; x[i] = x[r]; x[r] = t; }
    }
}
#define MIN_NUM_CHUNKS_FOR_COMPACTION ((size_t)1 << 7)   /* domain.rs:776-778 */
/* io_helper_with_roots (DIF, in-order in → bit-reversed out) — domain.rs:691-735 ;
 * butterfly_fn_io :651-656 ; apply_butterfly :667-688 */
static void io_helper(fr_t *x, size_t n, const fr_t *roots_in) {
    size_t nroots = n / 2;
    fr_t *roots = (fr_t *)malloc((nroots ? nroots : 1) * sizeof *roots);
    memcpy(roots, roots_in, nroots * sizeof *roots);
    size_t step = 1; int first = 1;
    for (size_t gap = n / 2; gap > 0; gap /= 2) {
        size_t chunk = 2 * gap, num_chunks = n / chunk;
        if (num_chunks >= MIN_NUM_CHUNKS_FOR_COMPACTION) {
            if (!first) { size_t m = 0; for (size_t i = 0; i < nroots; i += step * 2) roots[m++] = roots[i]; nroots = m; }
            step = 1;
        } else step = num_chunks;
        first = 0;
        const size_t st = step;
#pragma omp parallel for schedule(static)
        for (size_t b = 0; b < n / 2; b++) {
            size_t ch = b / gap, k = b % gap;
            fr_t *lo = &x[ch * chunk + k], *hi = lo + gap;
            fr_t neg; fr_sub(&neg, lo, hi); fr_add(lo, lo, hi); fr_mul(hi, &neg, &roots[k * st]);
        }
    }
    free(roots);
}
/* oi_helper_with_roots (DIT, bit-reversed in → in-order out) — domain.rs:737-773 ; butterfly_fn_oi :659-664 */
static void oi_helper(fr_t *x, size_t n, const fr_t *roots_cache) {
    size_t half = n / 2;
    size_t cmax = half / 2 < half / MIN_NUM_CHUNKS_FOR_COMPACTION ? half / 2 : half / MIN_NUM_CHUNKS_FOR_COMPACTION;
    fr_t *compacted = (fr_t *)malloc((cmax ? cmax : 1) * sizeof *compacted);
    for (size_t gap = 1; gap < n; gap *= 2) {
        size_t chunk = 2 * gap, num_chunks = n / chunk;
        const fr_t *roots; size_t st;
        if (num_chunks >= MIN_NUM_CHUNKS_FOR_COMPACTION && gap < n / 2) {
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < gap; i++) compacted[i] = roots_cache[i * num_chunks];
            roots = compacted; st = 1;
        } else { roots = roots_cache; st = num_chunks; }
#pragma omp parallel for schedule(static)
        for (size_t b = 0; b < n / 2; b++) {
            size_t ch = b / gap, k = b % gap;
            fr_t *lo = &x[ch * chunk + k], *hi = lo + gap;
            fr_t t; fr_mul(&t, hi, &roots[k * st]);
            fr_sub(hi, lo, &t); fr_add(lo, lo, &t);
        }
    }
    free(compacted);
}
/* distribute_powers_and_mul_by_const — domain.rs:239-254 */
static void distribute_powers(fr_t *x, size_t n, const fr_t *g, const fr_t *c) {
    size_t chunk = 1024;
#pragma omp parallel for schedule(static)
    for (size_t s = 0; s < n; s += chunk) {
        fr_t p = fr_pow_u64(g, s); fr_mul(&p, &p, c);
        size_t e = s + chunk < n ? s + chunk : n;
        for (size_t i = s; i < e; i++) { fr_mul(&x[i], &x[i], &p); fr_mul(&p, &p, g); }
    }
}
/* in_order_fft_in_place — domain.rs:374-392 → fft_helper_in_place_with_pc(II) :537-557 */
static void fft_in_place(fr_t *x, unsigned lg) {
    size_t n = (size_t)1 << lg;
    if (n == 1) return;
    fr_t w = fr_root_of_unity(lg);
    fr_t *roots = roots_of_unity(&w, n / 2);
    io_helper(x, n, roots);
    derange(x, lg);
    free(roots);
}
/* in_order_ifft_in_place — :403-422 (derange → oi_helper → × size_inv) ; coset variant :424-444 */
static void ifft_in_place(fr_t *x, unsigned lg, int coset) {
    size_t n = (size_t)1 << lg;
    fr_t size_inv = fr_from_u64(n); fr_inverse(&size_inv, &size_inv);
    if (n > 1) {
        fr_t w = fr_root_of_unity(lg), wi; fr_inverse(&wi, &w);
        fr_t *roots = roots_of_unity(&wi, n / 2);
        derange(x, lg);
        oi_helper(x, n, roots);
        free(roots);
    }
    if (coset) {
        fr_t g, gi; memcpy(g.l, FR_GEN, 32); fr_inverse(&gi, &g);
        distribute_powers(x, n, &gi, &size_inv);
    } else {
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; i++) fr_mul(&x[i], &x[i], &size_inv);
    }
}

/* ------------------------------------------------------------------ */
/* Exported C entry points (ctypes)                                   */
/* ------------------------------------------------------------------ */
#define API __attribute__((visibility("default")))

API int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
API void oracle_set_num_threads(int t) {
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}
API void oracle_fr_mul(uint64_t *r, const uint64_t *a, const uint64_t *b) { fr_mul((fr_t *)r, (const fr_t *)a, (const fr_t *)b); }
API void oracle_fr_add(uint64_t *r, const uint64_t *a, const uint64_t *b) { fr_add((fr_t *)r, (const fr_t *)a, (const fr_t *)b); }
API void oracle_fr_sub(uint64_t *r, const uint64_t *a, const uint64_t *b) { fr_sub((fr_t *)r, (const fr_t *)a, (const fr_t *)b); }
API int oracle_fr_inverse(uint64_t *r, const uint64_t *a) { return fr_inverse((fr_t *)r, (const fr_t *)a); }
API void oracle_fq_mul(uint64_t *r, const uint64_t *a, const uint64_t *b) { fq_mul((fq_t *)r, (const fq_t *)a, (const fq_t *)b); }
API void oracle_fq_add(uint64_t *r, const uint64_t *a, const uint64_t *b) { fq_add((fq_t *)r, (const fq_t *)a, (const fq_t *)b); }
API void oracle_fq_sub(uint64_t *r, const uint64_t *a, const uint64_t *b) { fq_sub((fq_t *)r, (const fq_t *)a, (const fq_t *)b); }
API int oracle_fq_inverse(uint64_t *r, const uint64_t *a) { return fq_inverse((fq_t *)r, (const fq_t *)a); }

/* to_bigint / from_bigint — fp_256.rs:362-413 : canonical <-> Montgomery, n elements */
API void oracle_fr_from_mont(uint64_t *out, const uint64_t *in, size_t n) {
    fr_t one = {{1, 0, 0, 0}};
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) fr_mul((fr_t *)(out + 4 * i), (const fr_t *)(in + 4 * i), &one);
}
API void oracle_fr_to_mont(uint64_t *out, const uint64_t *in, size_t n) {
    fr_t r2; memcpy(r2.l, FR_R2, 32);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) fr_mul((fr_t *)(out + 4 * i), (const fr_t *)(in + 4 * i), &r2);
}
API void oracle_fq_to_mont(uint64_t *out, const uint64_t *in, size_t n) {
    fq_t r2; memcpy(r2.l, FQ_R2, 48);
    for (size_t i = 0; i < n; i++) fq_mul((fq_t *)(out + 6 * i), (const fq_t *)(in + 6 * i), &r2);
}
API void oracle_fq_from_mont(uint64_t *out, const uint64_t *in, size_t n) {
    fq_t one = {{1, 0, 0, 0, 0, 0}};
    for (size_t i = 0; i < n; i++) fq_mul((fq_t *)(out + 6 * i), (const fq_t *)(in + 6 * i), &one);
}


/* Σ a_i·b_i mod r for canonical (non-Montgomery) 4-limb inputs; result canonical.  Used by the
 * full-size MSM property test: Σ s_i·(k_i·G) = (Σ s_i·k_i)·G. */
API void oracle_fr_dot_canonical(uint64_t *out, const uint64_t *a, const uint64_t *b, size_t n) {
    fr_t total = {{0, 0, 0, 0}};
#pragma omp parallel
    {
        fr_t local = {{0, 0, 0, 0}};
#pragma omp for schedule(static)
        for (size_t i = 0; i < n; i++) { fr_t t; fr_mul(&t, (const fr_t *)(a + 4 * i), (const fr_t *)(b + 4 * i)); fr_add(&local, &local, &t); }
#pragma omp critical
        fr_add(&total, &total, &local);
    }
    fr_t r2; memcpy(r2.l, FR_R2, 32);
    fr_mul(&total, &total, &r2);            /* (Σ a·b·R^{-1})·R2·R^{-1} = Σ a·b */
    memcpy(out, total.l, 32);
}

/* Same contract as snarkvm_ntt (algorithms/cuda/src/lib.rs:42-49): in place, NN order only.
 * direction 0 = Forward, 1 = Inverse ; type 0 = Standard, 1 = Coset. */
API int oracle_ntt(uint64_t *inout, uint32_t lg, int order, int direction, int type) {
    if (order != 0 || lg > 47) return 1;
    fr_t *x = (fr_t *)inout; size_t n = (size_t)1 << lg;
    if (direction == 0) {
        if (type == 1) { fr_t g, one = fr_one(); memcpy(g.l, FR_GEN, 32); distribute_powers(x, n, &g, &one); }  /* coset_fft :201-206 */
        fft_in_place(x, lg);
    } else ifft_in_place(x, lg, type == 1);
    return 0;
}

/* PolyMultiplier::multiply — fft/polynomial/multiplier.rs:70-134, with the FFI shape of
 * snarkvm_polymul (cuda/src/lib.rs:51-60): out[2^lg] = iFFT( Π FFT(pad(poly_i)) · Π eval_j ). */
API int oracle_polymul(uint64_t *out, size_t pcount, const uint64_t *const *polys, const size_t *plens,
                       size_t ecount, const uint64_t *const *evals, const size_t *elens, uint32_t lg) {
    size_t n = (size_t)1 << lg;
    if (pcount + ecount == 0) return 0;
    fr_t *acc = (fr_t *)out, *tmp = (fr_t *)malloc(n * sizeof *tmp);
    int have = 0;
    for (size_t p = 0; p < pcount; p++) {
        if (plens[p] > n) { free(tmp); return 1; }
        fr_t *dst = have ? tmp : acc;
        memset(dst, 0, n * sizeof *dst); memcpy(dst, polys[p], plens[p] * 32);
        fft_in_place(dst, lg);
        if (have) {
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < n; i++) fr_mul(&acc[i], &acc[i], &tmp[i]);
        }
        have = 1;
    }
    for (size_t e = 0; e < ecount; e++) {
        if (elens[e] != n) { free(tmp); return 1; }
        const fr_t *src = (const fr_t *)evals[e];
        if (have) {
#pragma omp parallel for schedule(static)
            for (size_t i = 0; i < n; i++) fr_mul(&acc[i], &acc[i], &src[i]);
        } else memcpy(acc, src, n * sizeof *acc);
        have = 1;
    }
    ifft_in_place(acc, lg, 0);
    free(tmp);
    return 0;
}

/* ------------------------------------------------------------------ */
/* Next-row oracles (SURVEY §8 f1–f4)                                  */
/* ------------------------------------------------------------------ */
/* Projective · Fr.  The reference routes `Projective * ScalarField` (projective.rs:488-504) through the curve's
 * mul_projective (GLV for BLS12-377 G1); the group element is the same whichever ladder computes it, and every
 * comparison is made on the affine image, so this is plain MSB-first double-and-add over scalar.to_bigint(). */
static g1_proj_t proj_mul_fr(const g1_proj_t *p, const fr_t *scalar_mont) {
    fr_t one = {{1, 0, 0, 0}}, s; fr_mul(&s, scalar_mont, &one);
    g1_proj_t out = proj_zero();
    int started = 0;
    for (int bit = 255; bit >= 0; bit--) {
        int b = (s.l[bit >> 6] >> (bit & 63)) & 1;
        if (!started) { if (!b) continue; started = 1; }
        proj_double(&out);
        if (b) proj_add(&out, p);
    }
    return out;
}
static void proj_neg(g1_proj_t *p) { if (!fq_is_zero(&p->y)) { fq_t z = fq_zero(); fq_sub(&p->y, &z, &p->y); } }

/* UniversalParams::lagrange_basis — polycommit/kzg10/data_structures.rs:68-72:
 *   domain.ifft(powers_of_beta_g[0..n].to_projective()) then batch_normalization_into_affine.
 * ifft over T = G1Projective is the generic in_order_ifft_in_place (domain.rs:403-422; the GPU branch needs
 * size_of::<T>() == 32): derange → oi_helper (butterfly_fn_oi :659-664: hi *= root; neg = lo − hi; lo += hi; hi = neg)
 * → every value *= size_inv. */
API int oracle_g1_ifft(void *out104, const void *in104, uint32_t lg) {
    if (lg > 30) return 1;
    size_t n = (size_t)1 << lg;
    g1_proj_t *x = (g1_proj_t *)malloc(n * sizeof *x);
    for (size_t i = 0; i < n; i++) { g1_affine_t a; memcpy(&a, (const uint8_t *)in104 + i * 104, 104); a.inf = a.inf != 0; x[i] = aff_to_proj(&a); }
    fr_t size_inv = fr_from_u64(n); fr_inverse(&size_inv, &size_inv);
    if (n > 1) {
        fr_t w = fr_root_of_unity(lg), wi; fr_inverse(&wi, &w);
        fr_t *roots = roots_of_unity(&wi, n / 2);
        for (size_t i = 1; i + 1 < n; i++) {                 /* derange */
            size_t r = 0, v = i;
            for (unsigned b = 0; b < lg; b++) { r = (r << 1) | (v & 1); v >>= 1; }
            if (i < r) { g1_proj_t t = x[i]; x[i] = x[r]; x[r] = t; }
        }
        for (size_t gap = 1; gap < n; gap *= 2) {
            size_t chunk = 2 * gap, num_chunks = n / chunk;
#pragma omp parallel for schedule(dynamic, 16)
            for (size_t b = 0; b < n / 2; b++) {
                size_t ch = b / gap, k = b % gap;
                g1_proj_t *lo = &x[ch * chunk + k], *hi = lo + gap;
                g1_proj_t t = proj_mul_fr(hi, &roots[k * num_chunks]);
                g1_proj_t neg = t; proj_neg(&neg);
                g1_proj_t d = *lo; proj_add(&d, &neg);
                proj_add(lo, &t);
                *hi = d;
            }
        }
        free(roots);
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t i = 0; i < n; i++) {
        g1_proj_t v = proj_mul_fr(&x[i], &size_inv);
        g1_affine_t a = proj_to_affine(&v);
        memcpy((uint8_t *)out104 + i * 104, &a, 104);
    }
    free(x);
    return 0;
}

/* serial_batch_inversion_and_mul — fields/src/lib.rs:93-129: v_i ← coeff · v_i^{-1}, zeros stay zero */
API void oracle_fr_batch_inversion_and_mul(uint64_t *v, size_t n, const uint64_t *coeff) {
    fr_t *x = (fr_t *)v;
    fr_t *prod = (fr_t *)malloc((n ? n : 1) * sizeof *prod);
    size_t m = 0;
    fr_t tmp = fr_one();
    for (size_t i = 0; i < n; i++) if (!fr_is_zero(&x[i])) { fr_mul(&tmp, &tmp, &x[i]); prod[m++] = tmp; }
    fr_inverse(&tmp, &tmp);
    fr_mul(&tmp, &tmp, (const fr_t *)coeff);
    for (size_t i = n; i-- > 0;) {
        if (fr_is_zero(&x[i])) continue;
        m--;
        fr_t s = m ? prod[m - 1] : fr_one(), new_tmp;
        fr_mul(&new_tmp, &tmp, &x[i]);
        fr_mul(&x[i], &tmp, &s);
        tmp = new_tmp;
    }
    free(prod);
}

/* DensePolynomial::divide_by_vanishing_poly — fft/polynomial/dense.rs:162-169 → divide_with_q_and_r
 * (fft/polynomial/mod.rs:222-256) with the sparse divisor x^n − 1 (domain.rs vanishing_polynomial).
 * p has m coefficients (trailing zeros allowed); q gets max(m − n, 0) and r gets min(m, n) coefficient slots
 * (the caller trims leading-zero high coefficients as DensePolynomial does).  Returns the quotient length. */
API size_t oracle_poly_divide_by_vanishing(uint64_t *q_out, uint64_t *r_out, const uint64_t *p, size_t m, size_t n) {
    fr_t *rem = (fr_t *)malloc((m ? m : 1) * sizeof *rem);
    memcpy(rem, p, m * 32);
    size_t len = m;
    while (len && fr_is_zero(&rem[len - 1])) len--;
    size_t qlen = m > n ? m - n : 0;
    fr_t *q = (fr_t *)q_out;
    memset(q, 0, qlen * 32);
    fr_t minus_one = fr_one(), zero = {{0, 0, 0, 0}}; fr_sub(&minus_one, &zero, &minus_one);
    while (len && len - 1 >= n) {                                         /* remainder.degree() >= divisor.degree() */
        fr_t cur = rem[len - 1];                                          /* × leading_coefficient^{-1} = 1 */
        size_t d = len - 1 - n;
        q[d] = cur;
        fr_t t; fr_mul(&t, &cur, &minus_one); fr_sub(&rem[d], &rem[d], &t);   /* remainder[d + 0] −= cur·(−1) */
        fr_sub(&rem[d + n], &rem[d + n], &cur);                               /* remainder[d + n] −= cur·1 */
        while (len && fr_is_zero(&rem[len - 1])) len--;
    }
    size_t rl = m < n ? m : n;
    memset(r_out, 0, rl * 32);
    memcpy(r_out, rem, (len < rl ? len : rl) * 32);
    free(rem);
    return qlen;
}

/* DensePolynomial::evaluate — fft/polynomial/dense.rs:98-114: Σ c_i·z^i (Horner here; field results are canonical) */
API void oracle_poly_evaluate(uint64_t *out, const uint64_t *coeffs, size_t m, const uint64_t *point) {
    fr_t acc = {{0, 0, 0, 0}};
    for (size_t i = m; i-- > 0;) { fr_mul(&acc, &acc, (const fr_t *)point); fr_add(&acc, &acc, (const fr_t *)(coeffs + 4 * i)); }
    memcpy(out, acc.l, 32);
}

/* compute_witness_polynomial — polycommit/kzg10/mod.rs:220-241: `polynomial / &divisor` with divisor = x − point, i.e. the
 * quotient of divide_with_q_and_r (fft/polynomial/mod.rs:222-256) against the dense divisor [−point, 1]; the remainder p(point)
 * is dropped.  p has m coefficient slots; q_out gets max(m − 1, 0) slots (zero above the true degree).  Returns that length. */
API size_t oracle_poly_divide_by_linear(uint64_t *q_out, const uint64_t *p, size_t m, const uint64_t *point_mont) {
    size_t qslots = m ? m - 1 : 0;
    memset(q_out, 0, qslots * 32);
    fr_t *rem = (fr_t *)malloc((m ? m : 1) * sizeof *rem);
    memcpy(rem, p, m * 32);
    size_t len = m;
    while (len && fr_is_zero(&rem[len - 1])) len--;
    fr_t zero = {{0, 0, 0, 0}}, neg_point, lead_inv = fr_one();
    fr_sub(&neg_point, &zero, (const fr_t *)point_mont);
    fr_inverse(&lead_inv, &lead_inv);                                     /* divisor.leading_coefficient().inverse() = 1 */
    fr_t *q = (fr_t *)q_out;
    while (len && len - 1 >= 1) {
        fr_t cur; fr_mul(&cur, &rem[len - 1], &lead_inv);
        size_t d = len - 1 - 1;
        q[d] = cur;
        fr_t t; fr_mul(&t, &cur, &neg_point); fr_sub(&rem[d], &rem[d], &t);
        fr_t one = fr_one(); fr_mul(&t, &cur, &one); fr_sub(&rem[d + 1], &rem[d + 1], &t);
        while (len && fr_is_zero(&rem[len - 1])) len--;
    }
    free(rem);
    return qslots;
}

/* z_M = M·(public ‖ private) row by row — inner_product, snark/varuna/ahp/prover/round_functions/mod.rs:169-189
 * (called for A, B, C at :128-152).  CSR: row r owns entries [row_ptr[r], row_ptr[r+1]) of (vals, cols). */
API void oracle_sparse_matvec(uint64_t *out, const uint32_t *row_ptr, const uint32_t *cols, const uint64_t *vals, size_t nrows,
                              const uint64_t *public_vars, size_t num_public, const uint64_t *private_vars) {
    fr_t one = fr_one();
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < nrows; r++) {
        fr_t result = {{0, 0, 0, 0}};
        for (uint32_t e = row_ptr[r]; e < row_ptr[r + 1]; e++) {
            size_t i = cols[e];
            const fr_t *variable = i < num_public ? (const fr_t *)(public_vars + 4 * i) : (const fr_t *)(private_vars + 4 * (i - num_public));
            const fr_t *coefficient = (const fr_t *)(vals + 4 * e);
            fr_t t;
            if (memcmp(coefficient, &one, 32) == 0) t = *variable; else fr_mul(&t, variable, coefficient);
            fr_add(&result, &result, &t);
        }
        memcpy(out + 4 * r, &result, 32);
    }
}

/* normalise: p.to_affine().to_projective() — the byte image the parity tests compare */
static void write_normalised(uint64_t *out144, const g1_proj_t *p) {
    g1_affine_t a = proj_to_affine(p); g1_proj_t q = aff_to_proj(&a); memcpy(out144, &q, 144);
}
/* algo: 0 = VariableBase::msm G1 path (batched::msm, variable_base/mod.rs:30-49),
 *       1 = standard::msm, 2 = msm_naive.  raw != 0 ⇒ write the un-normalised Jacobian. */
API int oracle_msm(uint64_t *out144, const void *points, size_t npoints, const uint64_t *scalars, size_t stride,
                   int algo, int raw) {
    if (stride < 97) return 1;
    g1_proj_t r;
    if (algo == 0) r = msm_batched((const uint8_t *)points, stride, npoints, scalars, npoints);
    else if (algo == 1) r = msm_standard((const uint8_t *)points, stride, scalars, npoints);
    else r = msm_naive((const uint8_t *)points, stride, scalars, npoints);
    if (raw) memcpy(out144, &r, 144); else write_normalised(out144, &r);
    return 0;
}
API void oracle_g1_normalise(uint64_t *out144, const uint64_t *in144) { g1_proj_t p; memcpy(&p, in144, 144); write_normalised(out144, &p); }
API void oracle_g1_mul(uint64_t *out144, const void *affine104, const uint64_t *scalar) {
    g1_affine_t a; memcpy(&a, affine104, 104); a.inf = a.inf != 0;
    g1_proj_t r = aff_mul_bits(&a, scalar); write_normalised(out144, &r);
}
API void oracle_g1_add(uint64_t *out144, const uint64_t *a144, const uint64_t *b144) {
    g1_proj_t a, b; memcpy(&a, a144, 144); memcpy(&b, b144, 144); proj_add(&a, &b); write_normalised(out144, &a);
}
API int oracle_g1_is_on_curve(const void *affine104) {
    g1_affine_t a; memcpy(&a, affine104, 104);
    if (a.inf) return 1;
    fq_t y2, x3, one = fq_one(); fq_sqr(&y2, &a.y); fq_sqr(&x3, &a.x); fq_mul(&x3, &x3, &a.x); fq_add(&x3, &x3, &one);
    return fq_eq(&y2, &x3);
}
/* Pairwise affine sum through the batched-affine kernels (loop_1 → one inversion → loop_2):
 * out[i] = a[i] + b[i].  Exercises batch_add_loop_1/2 directly for the tests. */
API void oracle_batch_affine_add(void *out, const void *a, const void *b, size_t n) {
    g1_affine_t *A = (g1_affine_t *)malloc((n ? n : 1) * sizeof *A), *B = (g1_affine_t *)malloc((n ? n : 1) * sizeof *B);
    memcpy(A, a, n * 104); memcpy(B, b, n * 104);
    fq_t inv = fq_one(), half = fq_half();
    for (size_t i = 0; i < n; i++) { A[i].inf = A[i].inf != 0; B[i].inf = B[i].inf != 0; batch_add_loop_1(&A[i], &B[i], &half, &inv); }
    fq_inverse(&inv, &inv);
    for (size_t i = n; i-- > 0;) batch_add_loop_2(&A[i], &B[i], &inv);
    for (size_t i = 0; i < n; i++) if (A[i].inf) A[i] = aff_zero();
    memcpy(out, A, n * 104); free(A); free(B);
}

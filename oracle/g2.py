"""TEST INFRASTRUCTURE — Python big-int oracle for BLS12-377 G2 (points over Fq2) and the reference's generic Pippenger.

Restates, as plain mathematics on Python integers:
  * Fq2 = Fq[u]/(u² + 5)            — curves/src/bls12_377/fq2.rs:29-65 (NONRESIDUE = −5), fields/src/fp2.rs (mul / square / inverse)
  * G2: y² = x³ + B', B' = (0, b1)   — curves/src/bls12_377/g2.rs:38-107 (WEIERSTRASS_B, generator :228-282, cofactor)
  * in-memory layouts                — Affine<G2> = x.c0 x.c1 y.c0 y.c1 (4 × 48 B Montgomery) infinity pad: 200-byte stride;
                                       Projective<G2> = X Y Z (3 × 96 B): 288 bytes, zero = (0, 1, 0)
                                       (short_weierstrass_jacobian/{affine.rs:41-46, projective.rs:36-41,51-54})
  * standard::msm                    — algorithms/src/msm/variable_base/standard.rs:24-118 (the Pippenger every curve other than
                                       BLS12-377 G1 takes, msm/variable_base/mod.rs:30-49): unit scalars first, 2^c − 1 buckets
                                       per window, running sums, Horner over the windows; c = ln_without_floats(n) + 2.

Parity status: the constants are pinned to the reference's own numbers in tests/test_oracle_golden.py (NONRESIDUE, B', generator
limbs; generator on the curve and of order r — curves/src/bls12_377/tests.rs:673-678); there is no reference-held G2 MSM vector
(the reference tests the MSM against its own naive sum, msm/variable_base/mod.rs:90-119), so the MSM restatement is checked
against the naive double-and-add sum here.  Only tests/ may import this module.
"""
from __future__ import annotations

from .bls12_377 import FQ_MONT_R, Q_MOD, R_MOD, fq_from_mont, fq_to_mont

NONRESIDUE = Q_MOD - 5
G2_B = (0, 155198655607781456406391640216936120121836107652948796323930557600032281009004493664981332883744016074664192874906)
# G2_GENERATOR_{X,Y}_{C0,C1} of curves/src/bls12_377/g2.rs:228-282, converted out of Montgomery form (tests/test_oracle_golden.py
# redoes the conversion from the limbs extracted by tests/golden/make_golden.py)
G2_GEN = (
    (170590608266080109581922461902299092015242589883741236963254737235977648828052995125541529645051927918098146183295,
     83407003718128594709087171351153471074446327721872642659202721143408712182996929763094113874399921859453255070254),
    (1843833842842620867708835993770650838640642469700861403869757682057607397502738488921663703124647238454792872005,
     33145532013610981697337930729788870077912093258611421158732879580766461459275194744385880708057348608045241477209),
)
G2_AFFINE_STRIDE = 200
G2_PROJECTIVE_BYTES = 288


# ---- Fq2 ----
def f2_add(a, b): return ((a[0] + b[0]) % Q_MOD, (a[1] + b[1]) % Q_MOD)
def f2_sub(a, b): return ((a[0] - b[0]) % Q_MOD, (a[1] - b[1]) % Q_MOD)
def f2_neg(a): return ((-a[0]) % Q_MOD, (-a[1]) % Q_MOD)
def f2_mul(a, b): return ((a[0] * b[0] + NONRESIDUE * a[1] * b[1]) % Q_MOD, (a[0] * b[1] + a[1] * b[0]) % Q_MOD)
def f2_sqr(a): return f2_mul(a, a)
def f2_inv(a):
    n = pow((a[0] * a[0] - NONRESIDUE * a[1] * a[1]) % Q_MOD, -1, Q_MOD)
    return (a[0] * n % Q_MOD, (-a[1]) * n % Q_MOD)
F2_ZERO, F2_ONE = (0, 0), (1, 0)


# ---- G2 (affine points as ((x0, x1), (y0, y1)) or None) ----
def g2_is_on_curve(p) -> bool:
    if p is None:
        return True
    x, y = p
    return f2_sqr(y) == f2_add(f2_mul(f2_sqr(x), x), G2_B)


def g2_neg(p):
    return None if p is None else (p[0], f2_neg(p[1]))


def g2_add(p, q):
    if p is None: return q
    if q is None: return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        if y1 != y2 or y1 == F2_ZERO:
            return None
        lam = f2_mul(f2_mul((3, 0), f2_sqr(x1)), f2_inv(f2_add(y1, y1)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_sqr(lam), x1), x2)
    return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1))


# Jacobian arithmetic for speed (X, Y, Z), Z = 0 at infinity
def _jdbl(P):
    X, Y, Z = P
    if Z == F2_ZERO: return P
    A, B = f2_sqr(X), f2_sqr(Y)
    C = f2_sqr(B)
    D = f2_sub(f2_sub(f2_sqr(f2_add(X, B)), A), C); D = f2_add(D, D)
    E = f2_add(f2_add(A, A), A)
    X3 = f2_sub(f2_sqr(E), f2_add(D, D))
    C8 = f2_add(C, C); C8 = f2_add(C8, C8); C8 = f2_add(C8, C8)
    Y3 = f2_sub(f2_mul(E, f2_sub(D, X3)), C8)
    Z3 = f2_mul(f2_add(Y, Y), Z)
    return (X3, Y3, Z3)


def _jadd_affine(P, q):
    if q is None: return P
    X1, Y1, Z1 = P
    if Z1 == F2_ZERO: return (q[0], q[1], F2_ONE)
    Z1Z1 = f2_sqr(Z1)
    U2, S2 = f2_mul(q[0], Z1Z1), f2_mul(f2_mul(q[1], Z1), Z1Z1)
    if U2 == X1:
        return _jdbl(P) if S2 == Y1 else (F2_ONE, F2_ONE, F2_ZERO)
    H, R = f2_sub(U2, X1), f2_sub(S2, Y1)
    HH = f2_sqr(H); HHH = f2_mul(H, HH); V = f2_mul(X1, HH)
    X3 = f2_sub(f2_sub(f2_sqr(R), HHH), f2_add(V, V))
    Y3 = f2_sub(f2_mul(R, f2_sub(V, X3)), f2_mul(Y1, HHH))
    return (X3, Y3, f2_mul(Z1, H))


def _jadd(P, Q):
    if Q[2] == F2_ZERO: return P
    if P[2] == F2_ZERO: return Q
    return _jadd_affine(P, _jaff(Q))


def _jaff(P):
    X, Y, Z = P
    if Z == F2_ZERO: return None
    zi = f2_inv(Z); zi2 = f2_sqr(zi)
    return (f2_mul(X, zi2), f2_mul(f2_mul(Y, zi2), zi))


J_INF = (F2_ONE, F2_ONE, F2_ZERO)


def g2_mul(p, k: int):
    k %= R_MOD
    acc = J_INF
    for bit in bin(k)[2:] if k else "":
        acc = _jdbl(acc)
        if bit == "1":
            acc = _jadd_affine(acc, p)
    return _jaff(acc)


def msm_naive(bases, scalars):
    acc = J_INF
    for p, s in zip(bases, scalars):
        acc = _jadd(acc, (lambda a: J_INF if a is None else (a[0], a[1], F2_ONE))(g2_mul(p, s)))
    return _jaff(acc)


# ---- standard::msm (standard.rs:24-118), restated on the Jacobian helpers above ----
def ln_without_floats(a: int) -> int:
    """msm/mod.rs: (log2(a) * 69 / 100)"""
    return (a.bit_length() - 1 if a & (a - 1) == 0 else a.bit_length()) * 69 // 100     # log2 rounds up for non powers of two


def standard_msm(bases, scalars):
    n = min(len(bases), len(scalars))
    bases, scalars = bases[:n], [s % (1 << 256) for s in scalars[:n]]
    c = 1 if n < 32 else ln_without_floats(n) + 2
    num_bits = 253
    sums = []
    for w_start in range(0, num_bits, c):
        res = J_INF
        if w_start == 0:
            for s, b in zip(scalars, bases):
                if s == 1: res = _jadd_affine(res, b)
        window_size = w_start % c if w_start % c else c
        buckets = [J_INF] * ((1 << window_size) - 1)
        for s, b in zip(scalars, bases):
            if s > 1:
                d = (s >> w_start) % (1 << c)
                if d: buckets[d - 1] = _jadd_affine(buckets[d - 1], b)
        running = J_INF
        for bk in reversed(buckets):
            running = _jadd(running, bk)
            res = _jadd(res, running)
        sums.append((res, window_size))
    lowest, rest = sums[0], sums[1:]
    total = J_INF
    for s, ws in reversed(rest):
        total = _jadd(total, s)
        for _ in range(ws): total = _jdbl(total)
    return _jaff(_jadd(total, lowest[0]))


# ---- layouts ----
def g2_affine_bytes(p) -> bytes:
    if p is None:
        x, y, inf = F2_ZERO, F2_ONE, 1                                       # Affine::zero() = (0, 1, true)
    else:
        x, y, inf = p[0], p[1], 0
    return b"".join(fq_to_mont(v).to_bytes(48, "little") for v in (x[0], x[1], y[0], y[1])) + bytes([inf]) + b"\0" * 7


def g2_affine_from_bytes(b: bytes):
    if b[192] != 0:
        return None
    v = [fq_from_mont(int.from_bytes(b[48 * i:48 * i + 48], "little")) for i in range(4)]
    return ((v[0], v[1]), (v[2], v[3]))


def g2_projective_bytes_normalised(p) -> bytes:
    """288-byte image of p.to_projective(): (x, y, 1) or (0, 1, 0)"""
    if p is None:
        coords = (F2_ZERO, F2_ONE, F2_ZERO)
    else:
        coords = (p[0], p[1], F2_ONE)
    return b"".join(fq_to_mont(c).to_bytes(48, "little") for f in coords for c in f)

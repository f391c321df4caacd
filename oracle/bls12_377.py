"""TEST INFRASTRUCTURE — Python big-int root of trust for the BLS12-377 hot path.

This module is the *oracle of oracles*: plain Python integers, no limbs, no
Montgomery tricks.  It is used only by ``tests/`` (and by ``oracle/`` self
checks) to pin the C restatement (``oracle/oracle.c``) and, through it, the
CUDA kernels.  Nothing under ``snarkvm_b200/`` may import it.

Every constant below is a mathematical fact restated from the reference
(numbers, not code):

* Fr  — curves/src/bls12_377/fr.rs:109-192   (modulus, R, R2, INV, GENERATOR=22,
        TWO_ADICITY=47, TWO_ADIC_ROOT_OF_UNITY :115-120)
* Fq  — curves/src/bls12_377/fq.rs:85-176    (modulus, R, R2, INV, TWO_ADICITY=46)
* G1  — curves/src/bls12_377/g1.rs:78-91,219-253 (y^2 = x^3 + 1, generator)

Algorithms restated:

* naive MSM (double-and-add)  — algorithms/src/msm/variable_base/mod.rs:52-57,
  curves/src/templates/short_weierstrass_jacobian/affine.rs:173-182 (`mul_bits`)
* radix-2 evaluation domain   — algorithms/src/fft/domain.rs:118-147 (`new`),
  :169-221 (fft / ifft / coset variants), fields/src/traits/fft_field.rs:38-86
  (`get_root_of_unity`)
* in-memory layouts           — fields/src/fp_256.rs:52, fp_384.rs:52 (LE u64
  Montgomery limbs), short_weierstrass_jacobian/affine.rs:41-46 (x, y, infinity;
  104-byte stride), projective.rs:36-41 (X, Y, Z; 144 bytes; zero = (0, R, 0))

Parity status: pinned against the reference's own constants / KATs in
tests/test_oracle_golden.py (root-of-unity table, domain elements of
circuit_0, G1 generator, real SRS points).
"""
from __future__ import annotations

import struct

# --------------------------------------------------------------------------
# Field parameters
# --------------------------------------------------------------------------
R_MOD = 8444461749428370424248824938781546531375899335154063827935233455917409239041  # Fr modulus
Q_MOD = 258664426012969094010652733694893533536393512754914660539884262666720468348340822774968888139573360124440321458177  # Fq modulus

FR_BITS = 253
FQ_BITS = 377
FR_LIMBS = 4
FQ_LIMBS = 6
FR_MONT_R = (1 << 256) % R_MOD
FQ_MONT_R = (1 << 384) % Q_MOD
FR_MONT_R2 = FR_MONT_R * FR_MONT_R % R_MOD
FQ_MONT_R2 = FQ_MONT_R * FQ_MONT_R % Q_MOD
FR_INV64 = (-pow(R_MOD, -1, 1 << 64)) % (1 << 64)
FQ_INV64 = (-pow(Q_MOD, -1, 1 << 64)) % (1 << 64)

FR_TWO_ADICITY = 47
FR_GENERATOR = 22
FR_TWO_ADIC_ROOT = 8065159656716812877374967518403273466521432693661810619979959746626482506078

G1_B = 1
G1_GEN_X = 89363714989903307245735717098563574705733591463163614225748337416674727625843187853442697973404985688481508350822
G1_GEN_Y = 3702177272937190650578065972808860481433820514072818216637796320125658674906330993856598323293086021583822603349

AFFINE_STRIDE = 104   # x[48] y[48] inf[1] pad[7]
PROJECTIVE_BYTES = 144


# --------------------------------------------------------------------------
# limb (de)serialisation
# --------------------------------------------------------------------------
def to_limbs(v: int, n: int) -> list[int]:
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def from_limbs(limbs) -> int:
    v = 0
    for i, l in enumerate(limbs):
        v |= int(l) << (64 * i)
    return v


def fr_to_mont(v: int) -> int:
    return v * FR_MONT_R % R_MOD


def fr_from_mont(m: int) -> int:
    return m * pow(FR_MONT_R, -1, R_MOD) % R_MOD


def fq_to_mont(v: int) -> int:
    return v * FQ_MONT_R % Q_MOD


def fq_from_mont(m: int) -> int:
    return m * pow(FQ_MONT_R, -1, Q_MOD) % Q_MOD


def fr_bytes_mont(v: int) -> bytes:
    """32-byte in-memory image of an `Fr` (Montgomery, LE)."""
    return fr_to_mont(v).to_bytes(32, "little")


def fr_from_bytes_mont(b: bytes) -> int:
    return fr_from_mont(int.from_bytes(b, "little"))


def affine_bytes(p) -> bytes:
    """104-byte in-memory image of `Affine<G1>`; p = None (infinity) or (x, y).

    Affine::zero() is (0, 1, true) — affine.rs:57-59."""
    if p is None:
        x, y, inf = 0, 1, 1
    else:
        x, y, inf = p[0], p[1], 0
    return (fq_to_mont(x).to_bytes(48, "little") + fq_to_mont(y).to_bytes(48, "little")
            + bytes([inf]) + b"\0" * 7)


def affine_from_bytes(b: bytes):
    if b[96] != 0:
        return None
    return (fq_from_mont(int.from_bytes(b[0:48], "little")),
            fq_from_mont(int.from_bytes(b[48:96], "little")))


def projective_bytes_normalised(p) -> bytes:
    """144-byte image of `p.to_affine().to_projective()` (Z = R, or (0, R, 0))."""
    if p is None:
        x, y, z = 0, 1, 0
    else:
        x, y, z = p[0], p[1], 1
    return b"".join(fq_to_mont(c).to_bytes(48, "little") for c in (x, y, z))


def projective_from_bytes(b: bytes):
    """Decode any Jacobian (X, Y, Z) image to affine ints (or None)."""
    X, Y, Z = (fq_from_mont(int.from_bytes(b[i:i + 48], "little")) for i in (0, 48, 96))
    if Z == 0:
        return None
    zi = pow(Z, -1, Q_MOD)
    return (X * zi * zi % Q_MOD, Y * zi * zi * zi % Q_MOD)


# --------------------------------------------------------------------------
# G1 group law on affine integer coordinates (None = infinity)
# --------------------------------------------------------------------------
def g1_is_on_curve(p) -> bool:
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - G1_B) % Q_MOD == 0


def g1_neg(p):
    return None if p is None else (p[0], (-p[1]) % Q_MOD)


def g1_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if (y1 + y2) % Q_MOD == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, Q_MOD) % Q_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q_MOD) % Q_MOD
    x3 = (lam * lam - x1 - x2) % Q_MOD
    y3 = (lam * (x1 - x3) - y1) % Q_MOD
    return (x3, y3)


# Jacobian arithmetic for speed in scalar multiplication (a = 0).
def _jac_double(P):
    X, Y, Z = P
    if Z == 0:
        return P
    A = X * X % Q_MOD
    B = Y * Y % Q_MOD
    C = B * B % Q_MOD
    D = 2 * ((X + B) * (X + B) - A - C) % Q_MOD
    E = 3 * A % Q_MOD
    F = E * E % Q_MOD
    X3 = (F - 2 * D) % Q_MOD
    Y3 = (E * (D - X3) - 8 * C) % Q_MOD
    Z3 = 2 * Y * Z % Q_MOD
    return (X3, Y3, Z3)


def _jac_add_affine(P, q):
    if q is None:
        return P
    X1, Y1, Z1 = P
    x2, y2 = q
    if Z1 == 0:
        return (x2, y2, 1)
    Z1Z1 = Z1 * Z1 % Q_MOD
    U2 = x2 * Z1Z1 % Q_MOD
    S2 = y2 * Z1 * Z1Z1 % Q_MOD
    if U2 == X1:
        if S2 == Y1:
            return _jac_double(P)
        return (0, 1, 0)
    H = (U2 - X1) % Q_MOD
    HH = H * H % Q_MOD
    HHH = H * HH % Q_MOD
    r = (S2 - Y1) % Q_MOD
    V = X1 * HH % Q_MOD
    X3 = (r * r - HHH - 2 * V) % Q_MOD
    Y3 = (r * (V - X3) - Y1 * HHH) % Q_MOD
    Z3 = Z1 * H % Q_MOD
    return (X3, Y3, Z3)


def _jac_to_affine(P):
    X, Y, Z = P
    if Z == 0:
        return None
    zi = pow(Z, -1, Q_MOD)
    zi2 = zi * zi % Q_MOD
    return (X * zi2 % Q_MOD, Y * zi2 * zi % Q_MOD)


def g1_mul(p, k: int):
    """k·p by MSB-first double-and-add (the reference's `mul_bits`)."""
    if p is None or k == 0:
        return None
    acc = (0, 1, 0)
    for bit in bin(k)[2:]:
        acc = _jac_double(acc)
        if bit == "1":
            acc = _jac_add_affine(acc, p)
    return _jac_to_affine(acc)


G1_GENERATOR = (G1_GEN_X, G1_GEN_Y)


def msm_naive(bases, scalars):
    """Σ scalars[i]·bases[i] over zip(bases, scalars) — variable_base/mod.rs:52-57."""
    acc = None
    for p, s in zip(bases, scalars):
        acc = g1_add(acc, g1_mul(p, s))
    return acc


# --------------------------------------------------------------------------
# Evaluation domain (textbook O(n log n) on plain integers mod r)
# --------------------------------------------------------------------------
def fr_root_of_unity(n: int) -> int:
    """Primitive n-th root (n a power of two ≤ 2^47): fft_field.rs:38-86."""
    log_n = n.bit_length() - 1
    assert 1 << log_n == n and log_n <= FR_TWO_ADICITY
    w = FR_TWO_ADIC_ROOT
    for _ in range(FR_TWO_ADICITY - log_n):
        w = w * w % R_MOD
    return w


def _ntt_rec(a, w):
    n = len(a)
    if n == 1:
        return a
    e = _ntt_rec(a[0::2], w * w % R_MOD)
    o = _ntt_rec(a[1::2], w * w % R_MOD)
    out = [0] * n
    t = 1
    for k in range(n // 2):
        x = o[k] * t % R_MOD
        out[k] = (e[k] + x) % R_MOD
        out[k + n // 2] = (e[k] - x) % R_MOD
        t = t * w % R_MOD
    return out


def fft(a):
    """y_k = Σ_j a_j ω^{jk}; natural order in and out (domain.rs:169-175)."""
    n = len(a)
    return _ntt_rec(list(a), fr_root_of_unity(n)) if n > 1 else list(a)


def ifft(a):
    n = len(a)
    if n == 1:
        return list(a)
    w_inv = pow(fr_root_of_unity(n), -1, R_MOD)
    n_inv = pow(n, -1, R_MOD)
    return [x * n_inv % R_MOD for x in _ntt_rec(list(a), w_inv)]


def coset_fft(a):
    """x_j ← x_j·g^j, then fft (domain.rs:201-206)."""
    g = 1
    b = []
    for x in a:
        b.append(x * g % R_MOD)
        g = g * FR_GENERATOR % R_MOD
    return fft(b)


def coset_ifft(a):
    """ifft, then y_j ← y_j·g^{-j} (domain.rs:424-444)."""
    gi = pow(FR_GENERATOR, -1, R_MOD)
    out = []
    g = 1
    for x in ifft(a):
        out.append(x * g % R_MOD)
        g = g * gi % R_MOD
    return out


def dft_horner(a):
    """O(n²) evaluation at every domain point (fft/tests.rs:119-149)."""
    n = len(a)
    w = fr_root_of_unity(n) if n > 1 else 1
    out = []
    x = 1
    for _ in range(n):
        acc = 0
        for c in reversed(a):
            acc = (acc * x + c) % R_MOD
        out.append(acc)
        x = x * w % R_MOD
    return out


# --------------------------------------------------------------------------
# .usrs parser (uncompressed canonical points) — used for the real-SRS fixture
# --------------------------------------------------------------------------
def parse_usrs_points(blob: bytes, count: int | None = None):
    """u64-LE count, then count × 96 B (x LE, y LE; flags in top bits of last byte:
    bit7 = y sign, bit6 = infinity) — utilities/src/serialize/flags.rs:72-98."""
    (n,) = struct.unpack_from("<Q", blob, 0)
    if count is not None:
        n = min(n, count)
    pts = []
    off = 8
    for _ in range(n):
        xb = blob[off:off + 48]
        yb = bytearray(blob[off + 48:off + 96])
        flags = yb[47] & 0xC0
        yb[47] &= 0x3F
        off += 96
        if flags & 0x40:
            pts.append(None)
        else:
            pts.append((int.from_bytes(xb, "little"), int.from_bytes(bytes(yb), "little")))
    return pts

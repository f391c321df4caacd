"""Statement-level model of ff.cuh's Karatsuba Montgomery multiplier (mul_wide E/O carry chains, subtractive
Karatsuba assembly, word-serial Montgomery reduction of the 2N-limb product), checked against big integers.
Run: python tests/manual/karatsuba_model.py"""
import random

M32 = 0xffffffff


class CC:
    """the PTX carry flag"""
    def __init__(self): self.c = 0
    def add_cc(self, a, b): s = a + b; self.c = s >> 32; return s & M32
    def addc_cc(self, a, b): s = a + b + self.c; self.c = s >> 32; return s & M32
    def addc(self, a, b): s = a + b + self.c; assert s >> 32 == 0, "carry lost"; return s & M32
    def addc_drop(self, a, b): s = a + b + self.c; return s & M32
    def sub_cc(self, a, b): s = a - b; self.c = 1 if s < 0 else 0; return s & M32          # c = borrow
    def subc_cc(self, a, b): s = a - b - self.c; self.c = 1 if s < 0 else 0; return s & M32
    def mad_lo_cc(self, a, b, c): s = ((a * b) & M32) + c; self.c = s >> 32; return s & M32
    def madc_lo_cc(self, a, b, c): s = ((a * b) & M32) + c + self.c; self.c = s >> 32; return s & M32
    def madc_hi_cc(self, a, b, c): s = ((a * b) >> 32) + c + self.c; self.c = s >> 32; return s & M32


def mul_wide(a, b, H):
    """T[2H] = a[H]·b[H] with even/odd position accumulators (E: even positions, O: odd positions, offset one limb)"""
    cc = CC()
    E = [0] * (2 * H); O = [0] * (2 * H)
    for i in range(H):
        first = True; p = None
        for j in range(i & 1, H, 2):
            p = i + j
            E[p] = cc.mad_lo_cc(a[j], b[i], E[p]) if first else cc.madc_lo_cc(a[j], b[i], E[p])
            E[p + 1] = cc.madc_hi_cc(a[j], b[i], E[p + 1])
            first = False
        if p is not None:
            if p + 2 < 2 * H: E[p + 2] = cc.addc(E[p + 2], 0)
            else: assert cc.c == 0
        first = True; p = None
        for j in range(1 - (i & 1), H, 2):
            p = i + j                                  # odd position → O[p-1], O[p]
            O[p - 1] = cc.mad_lo_cc(a[j], b[i], O[p - 1]) if first else cc.madc_lo_cc(a[j], b[i], O[p - 1])
            O[p] = cc.madc_hi_cc(a[j], b[i], O[p])
            first = False
        if p is not None:
            if p + 1 < 2 * H - 1: O[p + 1] = cc.addc(O[p + 1], 0)
            else: assert cc.c == 0
    T = [0] * (2 * H)
    T[0] = E[0]
    T[1] = cc.add_cc(E[1], O[0])
    for k in range(2, 2 * H): T[k] = cc.addc_cc(E[k], O[k - 1])
    assert cc.c == 0 and O[2 * H - 1] == 0
    return T


def limbs(v, n): return [(v >> (32 * i)) & M32 for i in range(n)]
def val(l): return sum(x << (32 * i) for i, x in enumerate(l))


def abs_diff(x, y, H):
    """|x − y| and sign (1 if x < y)"""
    cc = CC()
    d = [0] * H
    d[0] = cc.sub_cc(x[0], y[0])
    for i in range(1, H): d[i] = cc.subc_cc(d_i := x[i], y[i]) if False else cc.subc_cc(x[i], y[i])
    neg = cc.c
    if neg:                                             # two's complement
        m = M32
        d2 = [0] * H
        d2[0] = cc.add_cc(d[0] ^ m, 1)
        for i in range(1, H): d2[i] = cc.addc_cc(d[i] ^ m, 0)
        d = d2
    return d, neg


def karatsuba_wide(a, b, N):
    H = N // 2
    cc = CC()
    z0 = mul_wide(a[:H], b[:H], H)
    z2 = mul_wide(a[H:], b[H:], H)
    da, sa = abs_diff(a[:H], a[H:], H)
    db, sb = abs_diff(b[:H], b[H:], H)
    zm = mul_wide(da, db, H)
    # mid = z0 + z2 ∓ zm   (N + 1 limbs)
    mid = [0] * (N + 1)
    mid[0] = cc.add_cc(z0[0], z2[0])
    for k in range(1, N): mid[k] = cc.addc_cc(z0[k], z2[k])
    mid[N] = cc.addc(0, 0)
    mask = M32 if sa == sb else 0                       # subtract zm when the signs agree
    mid[0] = cc.add_cc(mid[0], zm[0] ^ mask)            # … + (zm ^ mask) + (mask & 1): the +1 goes in as a second chain below
    for k in range(1, N): mid[k] = cc.addc_cc(mid[k], zm[k] ^ mask)
    mid[N] = cc.addc_drop(mid[N], mask)
    if mask:
        mid[0] = cc.add_cc(mid[0], 1)
        for k in range(1, N + 1): mid[k] = cc.addc_cc(mid[k], 0) if k < N else cc.addc_drop(mid[k], 0)
    T = z0 + z2
    T[H] = cc.add_cc(T[H], mid[0])
    for k in range(1, N + 1): T[H + k] = cc.addc_cc(T[H + k], mid[k])
    for k in range(H + N + 1, 2 * N): T[k] = cc.addc_cc(T[k], 0)
    assert cc.c == 0
    return T


def mont_reduce_wide(T, p, N, inv32):
    """word-serial Montgomery reduction of a 2N-limb T < p·2^{32N}: returns T·2^{-32N} mod p (plain model: one limb per step)"""
    T = list(T) + [0]
    cc = CC()
    for i in range(N):
        m = (T[i] * inv32) & M32
        carry = 0
        for j in range(N):
            s = T[i + j] + m * p[j] + carry
            T[i + j] = s & M32; carry = s >> 32
        k = i + N
        while carry:
            s = T[k] + carry; T[k] = s & M32; carry = s >> 32; k += 1
    r = val(T[N:2 * N + 1])
    P = val(p)
    return r - P if r >= P else r


if __name__ == "__main__":
    Q = 0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001
    R = 8444461749428370424248824938781546531375899335154063827935233455917409239041
    rnd = random.Random(1)
    for P, N in ((Q, 12), (R, 8)):
        inv32 = (-pow(P, -1, 1 << 32)) & M32
        pl = limbs(P, N)
        cases = [(0, 0), (1, 1), (P - 1, P - 1), (P - 1, 1), ((1 << (32 * (N // 2))) - 1, P - 1), ((1 << (32 * (N // 2))), (1 << (32 * (N // 2))) - 1)]
        cases += [(rnd.randrange(P), rnd.randrange(P)) for _ in range(3000)]
        # halves equal / ordered both ways so that every sign combination occurs
        for _ in range(500):
            x = rnd.randrange(1 << (32 * (N // 2) - 8)); y = rnd.randrange(1 << (32 * (N // 2) - 8))
            cases.append(((x << (32 * (N // 2))) | y, (y << (32 * (N // 2))) | x))
            cases.append(((x << (32 * (N // 2))) | x, rnd.randrange(P)))
        for a, b in cases:
            a %= P; b %= P
            T = karatsuba_wide(limbs(a, N), limbs(b, N), N)
            assert val(T) == a * b, (hex(a), hex(b))
            assert val(mul_wide(limbs(a, N), limbs(b, N), N)) == a * b
            got = mont_reduce_wide(T, pl, N, inv32)
            assert got == a * b * pow(1 << (32 * N), -1, P) % P
        print("ok", N, len(cases))

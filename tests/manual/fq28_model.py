"""Exact Python model of experiments/fq28/fq28.cuh (radix-2^28, 14-limb, lazily reduced Fq arithmetic).

Mirrors the device code operation by operation on Python integers, asserting every 32-/64-bit register bound
the CUDA code relies on (column sums < 2^64, limbs < 2^32, borrow-proof subtraction), and checks the result
against plain modular arithmetic.  Run:  python tests/manual/fq28_model.py   (lives under tests/ because it uses the oracle; the CUDA side is experiments/fq28/)
"""
import random
import sys

sys.path.insert(0, __file__.rsplit("/tests/", 1)[0])
from oracle import bls12_377 as o  # noqa: E402

P = o.Q_MOD
B, N = 28, 14
MASK = (1 << B) - 1
RP = 1 << (B * N)              # R' = 2^392
U32, U64 = 1 << 32, 1 << 64


def limbs(v, raw_top=False):
    out = [(v >> (B * i)) & MASK for i in range(N)]
    if raw_top:
        out[N - 1] = v >> (B * (N - 1))
    return out


def value(l):
    return sum(x << (B * i) for i, x in enumerate(l))


P28 = limbs(P)


def borrowproof(K, OFF):
    c = limbs(K * P, raw_top=True)
    for j in range(N - 1):
        c[j] += 1 << OFF
        c[j + 1] -= 1 << (OFF - B)
    assert all(x >= 0 for x in c) and value(c) == K * P
    return c


C2P_28, C16P_28, C8P_30 = borrowproof(2, 28), borrowproof(16, 28), borrowproof(8, 30)


def chk32(l):
    assert all(0 <= x < U32 for x in l), [hex(x) for x in l]
    return l


def mul(a, b, square=False):
    """interleaved Montgomery product, window t[i..i+13]; returns normalised limbs, value < 2p"""
    chk32(a); chk32(b)
    t = [0] * 29
    for i in range(N):
        if square:
            t[2 * i] += a[i] * a[i]
            for j in range(i + 1, N):
                d = 2 * a[j]
                assert d < U32
                t[i + j] += a[i] * d
        else:
            for j in range(N):
                t[i + j] += a[j] * b[i]
        assert all(x < U64 for x in t)
        m = (-t[i]) & MASK                     # PINV28 = -1
        s = t[i] + m
        assert s < U64 and s & MASK == 0
        c = s >> B
        for j in range(1, N):
            t[i + j] += P28[j] * m
        t[i + 1] += c
        assert all(x < U64 for x in t)
    out = []
    for k in range(N, 2 * N):
        out.append(t[k] & MASK)
        t[k + 1] += t[k] >> B
        assert t[k + 1] < U64
    assert t[2 * N] == 0 or True
    # top limb keeps everything above (value < 2p fits)
    out[N - 1] = out[N - 1] | ((t[2 * N] << B))
    return chk32(out)


def add(a, b):
    return chk32([x + y for x, y in zip(a, b)])


def sub(a, b, C):
    """a + C - b, limb-wise, C a borrow-proof multiple of p with C_j >= b_j"""
    assert all(c >= y for c, y in zip(C, b)), "borrow"
    return chk32([x + c - y for x, c, y in zip(a, C, b)])


def norm(a):
    out, carry = [], 0
    for j in range(N):
        s = a[j] + carry
        assert s < U64
        if j < N - 1:
            out.append(s & MASK); carry = s >> B
        else:
            out.append(s)
    return chk32(out)


def to_int(a):
    """the field element represented (internal form x·R' mod p)"""
    return value(a) * pow(RP, -1, P) % P


def from_int(x):
    return limbs(x * RP % P)


def madd(acc, q, negate):
    """XYZZ += affine (madd-2008-s), the exact op sequence of XYZZ28::add_affine's main path"""
    X, Y, ZZ, ZZZ = acc
    qx, qy = q
    if negate:
        qy = sub([0] * N, qy, C2P_28)
    U2 = mul(qx, ZZ); S2 = mul(qy, ZZZ)
    Pp = sub(U2, X, C16P_28)
    R = sub(S2, Y, C16P_28)
    PP = mul(Pp, Pp, square=True)
    PPP = mul(Pp, PP)
    Q = mul(X, PP)
    s = add(add(PPP, Q), Q)
    X3 = norm(sub(mul(R, R, square=True), s, C8P_30))
    t = sub(Q, X3, C16P_28)
    Y3 = norm(sub(mul(R, t), mul(Y, PPP), C2P_28))
    return [X3, Y3, mul(ZZ, PP), mul(ZZZ, PPP)]


def main():
    rng = random.Random(1)
    # --- mul / sqr correctness incl. lazy inputs at the documented bounds ---
    for it in range(300):
        x, y = rng.randrange(P), rng.randrange(P)
        a, b = from_int(x), from_int(y)
        r = mul(a, b)
        assert to_int(r) == x * y % P and value(r) < 2 * P
        assert to_int(mul(a, a, square=True)) == x * x % P
        la = sub(a, from_int(rng.randrange(P)), C16P_28)      # limbs < 2^29.6, value < 18p
        lb = sub(b, from_int(rng.randrange(P)), C16P_28)
        assert to_int(mul(la, lb)) == to_int(la) * to_int(lb) % P
        assert to_int(mul(la, la, square=True)) == to_int(la) ** 2 % P
    # worst-case limbs: all limbs at 2^30 - 1 for mul, 2^29.6 for sqr
    big = [(1 << 30) - 1] * (N - 1) + [1 << 14]
    mul(big, big)
    sq = [int(2 ** 29.6)] * (N - 1) + [1 << 14]
    mul(sq, sq, square=True)
    # --- long madd chains vs affine arithmetic ---
    G = o.G1_GENERATOR
    for trial in range(4):
        pts = [o.g1_mul(G, rng.randrange(1, 1 << 64)) for _ in range(40)]
        first = pts[0]
        acc = [from_int(first[0]), from_int(first[1]), from_int(1), from_int(1)]
        expect = first
        for q in pts[1:]:
            neg = rng.random() < 0.5
            acc = madd(acc, (from_int(q[0]), from_int(q[1])), neg)
            expect = o.g1_add(expect, o.g1_neg(q) if neg else q)
            # stored invariants of the device code
            assert all(l < (1 << 28) for l in acc[0][:-1]) and value(acc[0]) < 10 * P
            assert all(l < (1 << 28) for l in acc[1][:-1]) and value(acc[1]) < 4 * P
            assert value(acc[2]) < 2 * P and value(acc[3]) < 2 * P
        Xv, Yv, ZZv, ZZZv = (to_int(c) for c in acc)
        assert (Xv * pow(ZZv, -1, P) % P, Yv * pow(ZZZv, -1, P) % P) == expect
    print("fq28 model ok")


if __name__ == "__main__":
    main()

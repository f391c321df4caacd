"""Lane-level model of the warp-cooperative Montgomery multiplication in csrc/ff.cuh (coop_mul): limb j of an Fq element lives in
lane j (lanes 12..31 hold zeros), carry-save accumulation with one limb shift per iteration, ballot-based carry / borrow
resolution at the end.  Every statement mirrors one CUDA statement; run it to check the algorithm against big integers:
    python tests/manual/coop_mul_model.py"""
import random

Q = 258664426012969094010652733694893533536393512754914660539884262666720468348340822774968888139573360124440321458177
M32 = (1 << 32) - 1
N = 12
INV32 = (-pow(Q, -1, 1 << 32)) % (1 << 32)
P = [(Q >> (32 * j)) & M32 if j < N else 0 for j in range(32)]


def shfl(v, src):
    return [v[src] for _ in range(32)]


def shfl_down1(v):
    return [v[j + 1] if j + 1 < 32 else v[j] for j in range(32)]


def shfl_up1(v):
    return [v[j - 1] if j >= 1 else v[j] for j in range(32)]


def ballot(pred):
    return sum(1 << j for j in range(32) if pred[j])


def carries_in(G, Pm):
    """bit j = carry INTO lane j when lane k generates (G bit k) or propagates (P bit k); 32-bit wrap like the hardware"""
    A, B = (G | Pm) & M32, G & M32
    return (((A + B) & M32) ^ A ^ B) & M32


def coop_mul(a, b):
    lo, hi, ex = [0] * 32, [0] * 32, [0] * 32
    for i in range(N):
        bi = shfl(b, i)
        for j in range(32):                                   # V += a·b_i
            t = lo[j] + hi[j] * (1 << 32) + ex[j] * (1 << 64) + a[j] * bi[j]
            lo[j], hi[j], ex[j] = t & M32, (t >> 32) & M32, t >> 64
        m = shfl([(lo[j] * INV32) & M32 for j in range(32)], 0)
        for j in range(32):                                   # V += m·p
            t = lo[j] + hi[j] * (1 << 32) + ex[j] * (1 << 64) + m[j] * P[j]
            lo[j], hi[j], ex[j] = t & M32, (t >> 32) & M32, t >> 64
            assert ex[j] < 4
        assert lo[0] == 0
        t = shfl_down1(lo)
        t[31] = 0                                             # CUDA: lane 31 reads its own value; its lo is always 0 anyway
        for j in range(32):                                   # one limb down: V'_j = (V_j >> 32) + lo_{j+1}
            s = hi[j] + t[j]
            lo[j], hi[j], ex[j] = s & M32, ex[j] + (s >> 32), 0
    up = shfl_up1(hi)
    up[0] = 0
    s = [(lo[j] + up[j]) & M32 for j in range(32)]
    c = [(lo[j] + up[j]) >> 32 for j in range(32)]
    assert hi[N - 1] == 0 and all(x in (0, 1) for x in c)
    cin = carries_in(ballot([x == 1 for x in c]), ballot([x == M32 for x in s]))
    t = [(s[j] + ((cin >> j) & 1)) & M32 for j in range(32)]
    gt, lt = ballot([t[j] > P[j] for j in range(32)]), ballot([t[j] < P[j] for j in range(32)])
    if gt >= lt:                                              # t ≥ p (top differing limb decides; equal ⇒ subtract too)
        eq = ballot([t[j] == P[j] for j in range(32)])
        bin_ = carries_in(lt, eq)
        t = [(t[j] - P[j] - ((bin_ >> j) & 1)) & M32 for j in range(32)]
    return t


def to_lanes(v):
    return [(v >> (32 * j)) & M32 if j < N else 0 for j in range(32)]


def from_lanes(l):
    assert all(x == 0 for x in l[N:])
    return sum(l[j] << (32 * j) for j in range(N))


if __name__ == "__main__":
    rng = random.Random(1)
    Rinv = pow(1 << 384, -1, Q)
    cases = [(0, 0), (1, 1), (Q - 1, Q - 1), (Q - 1, 1), ((1 << 377) % Q, Q - 2)]
    cases += [(rng.randrange(Q), rng.randrange(Q)) for _ in range(3000)]
    # values that force long carry / borrow propagation chains
    cases += [((1 << (32 * k)) - 1, (1 << 384) % Q) for k in range(1, 12)]
    cases += [(Q - (1 << (32 * k)), rng.randrange(Q)) for k in range(1, 11)]
    for a, b in cases:
        got = from_lanes(coop_mul(to_lanes(a), to_lanes(b)))
        assert got == a * b * Rinv % Q, (a, b)
    # a Fermat inversion chain in Montgomery form
    x = rng.randrange(1, Q)
    xm = x * (1 << 384) % Q
    acc, started = None, False
    e = Q - 2
    base = to_lanes(xm)
    for bit in range(e.bit_length() - 1, -1, -1):
        if started:
            acc = coop_mul(acc, acc)
        if (e >> bit) & 1:
            acc = coop_mul(acc, base) if started else base
            started = True
    assert from_lanes(acc) == pow(x, -1, Q) * (1 << 384) % Q
    print("coop_mul model ok:", len(cases), "products + one inversion chain")

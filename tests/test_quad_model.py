"""CPU: a lane-level model of csrc/quad.cuh — four lanes share one XYZZ group operation, lane q computes the q-th product of every
multiplication step and the products are exchanged inside the quad — checked against the big-int group law of the oracle.  The model
mirrors the operand selection of quad_add / quad_dbl / quad_add_affine statement by statement (sel4 = quad_sel, the tuples are the
values quad_get broadcasts), so an edit of the step tables that breaks the algebra fails here without a GPU."""
import random

from oracle import bls12_377 as py

Q = py.Q_MOD


def _steps(a_ops, b_ops):
    """one multiplication step: lane q multiplies a_ops[q]·b_ops[q]; returns what each lane holds"""
    return [x * y % Q for x, y in zip(a_ops, b_ops)]


def quad_dbl(p):
    X, Y, ZZ, ZZZ = p
    if ZZ == 0:
        return p
    U = 2 * Y % Q
    r1 = _steps((U, X, U, X), (U, X, U, X)); V, XX = r1[0], r1[1]
    M = 3 * XX % Q
    r2 = _steps((U, X, M, V), (V, V, M, ZZ)); W, S, MM, ZZ3 = r2
    X3 = (MM - 2 * S) % Q
    r3 = _steps((M, W, W, W), ((S - X3) % Q, Y, ZZZ, ZZZ))
    return X3, (r3[0] - r3[1]) % Q, ZZ3, r3[2]


def quad_add(a, b):
    if b[2] == 0:
        return a
    if a[2] == 0:
        return b
    r1 = _steps((a[0], b[0], a[1], b[1]), (b[2], a[2], b[3], a[3])); U1, U2, S1, S2 = r1
    P, R = (U2 - U1) % Q, (S2 - S1) % Q
    if P == 0:
        return quad_dbl(a) if R == 0 else (0, 0, 0, 0)
    r2 = _steps((P, R, a[2], a[3]), (P, R, b[2], b[3])); PP, RR = r2[0], r2[1]
    r3 = _steps((P, U1, r2[2], P), (PP, PP, PP, PP)); PPP, Qv, ZZ3 = r3[0], r3[1], r3[2]
    assert r3[3] == PPP                                     # lane 3 recomputes PPP for its own step 4
    X3 = (RR - PPP - 2 * Qv) % Q
    r4 = _steps((R, S1, R, r2[3]), ((Qv - X3) % Q, PPP, PPP, PPP))
    return X3, (r4[0] - r4[1]) % Q, ZZ3, r4[3]


def quad_add_affine(a, p, negate):
    if p is None:
        return a
    x, y = p[0], (-p[1]) % Q if negate else p[1]
    if a[2] == 0:
        return x, y, 1, 1
    r1 = _steps((x, y, x, y), (a[2], a[3], a[2], a[3])); U2, S2 = r1[0], r1[1]
    P, R = (U2 - a[0]) % Q, (S2 - a[1]) % Q
    if P == 0:
        return (0, 0, 0, 0) if R != 0 else quad_dbl((x, y, 1, 1))
    r2 = _steps((P, R, P, R), (P, R, P, R)); PP, RR = r2[0], r2[1]
    r3 = _steps((P, a[0], a[2], P), (PP, PP, PP, PP)); PPP, Qv, ZZ3 = r3[0], r3[1], r3[2]
    X3 = (RR - PPP - 2 * Qv) % Q
    r4 = _steps((R, a[1], a[3], a[3]), ((Qv - X3) % Q, PPP, PPP, PPP))
    return X3, (r4[0] - r4[1]) % Q, ZZ3, r4[2]


def _affine(p):
    if p[2] == 0:
        return None
    return p[0] * pow(p[2], -1, Q) % Q, p[1] * pow(p[3], -1, Q) % Q


def _xyzz(pt, lam):
    """an XYZZ image of the affine point with ZZ = λ², ZZZ = λ³"""
    if pt is None:
        return 0, 0, 0, 0
    return pt[0] * lam * lam % Q, pt[1] * pow(lam, 3, Q) % Q, lam * lam % Q, pow(lam, 3, Q)


def test_quad_lane_formulas_match_the_group_law():
    rng = random.Random(11)
    pts = [py.g1_mul(py.G1_GENERATOR, rng.randrange(1, py.R_MOD)) for _ in range(6)]
    cases = [(pts[0], pts[1]), (pts[2], pts[2]), (pts[3], py.g1_neg(pts[3])), (None, pts[4]), (pts[5], None), (None, None)]
    for p, q in cases:
        for _ in range(3):
            a, b = _xyzz(p, rng.randrange(1, Q)), _xyzz(q, rng.randrange(1, Q))
            assert _affine(quad_add(a, b)) == py.g1_add(p, q)
            assert _affine(quad_dbl(a)) == py.g1_add(p, p)
            for negate in (False, True):
                want = py.g1_add(p, py.g1_neg(q) if (negate and q is not None) else q)
                assert _affine(quad_add_affine(a, q, negate)) == want
    # a chain, as the scans use it: Σ of six points in two association orders
    acc1 = (0, 0, 0, 0)
    for p in pts:
        acc1 = quad_add(acc1, _xyzz(p, rng.randrange(1, Q)))
    left = quad_add(quad_add(_xyzz(pts[0], 3), _xyzz(pts[1], 5)), _xyzz(pts[2], 7))
    right = quad_add(_xyzz(pts[3], 2), quad_add(_xyzz(pts[4], 9), _xyzz(pts[5], 11)))
    want = None
    for p in pts:
        want = py.g1_add(want, p)
    assert _affine(acc1) == _affine(quad_add(left, right)) == want


def _fold_level(entries, lgw):
    """k_combine_level_quad / _quadseq / one level of k_window_combine_quad: groups of 8 (acc, run) entries spanning 2^lgw buckets
    each → acc' = Σ acc_s + 2^lgw·Σ s·run_s, run' = Σ run_s (the last group may be short)"""
    out = []
    for g in range(0, len(entries), 8):
        grp = entries[g:g + 8]
        acc = sum(a for a, _ in grp) + (1 << lgw) * sum(s * r for s, (_, r) in enumerate(grp))
        out.append((acc, sum(r for _, r in grp)))
    return out


def _tail(buckets, chunk):
    """the reduction tail as msm_core launches it: per-chunk (Σ (b − lo + 1)·S_b, Σ S_b) pairs (k_bucket_reduce_quad: chunk = 8,
    k_bucket_reduce<PAIRS>: chunk = 16), 8:1 levels while more than 64 entries are left, then the in-CTA levels down to one"""
    entries = []
    for lo in range(0, len(buckets), chunk):
        part = buckets[lo:lo + chunk]
        entries.append((sum((i + 1) * s for i, s in enumerate(part)), sum(part)))
    lgw = chunk.bit_length() - 1
    launches = 0
    while len(entries) > 64:
        entries = _fold_level(entries, lgw); lgw += 3; launches += 1
    while True:                                             # k_window_combine_quad: 64 → 8 → 1 through shared memory
        entries = _fold_level(entries, lgw); lgw += 3
        if len(entries) == 1:
            break
    return entries[0][0], launches


def test_weighted_fold_levels_compute_the_bucket_sum():
    """Σ_b (b + 1)·S_b (the running-sum reduction of batched.rs:356-361) through chunk pairs and 8:1 weighted folds, with integers
    standing in for the bucket sums (the map S ↦ k·S is linear, so the identity carries over to group elements)"""
    rng = random.Random(3)
    for nb, chunk in ((8, 8), (16, 8), (128, 8), (1024, 8), (4, 8), (16384, 16), (65536, 16), (1 << 19, 16)):
        buckets = [rng.randrange(1 << 40) for _ in range(nb)]
        got, launches = _tail(buckets, chunk)
        assert got == sum((b + 1) * s for b, s in enumerate(buckets)), (nb, chunk)
        # the launch counts the host code relies on: 1024 buckets / 8 = 128 entries → one level kernel, 65536 / 16 = 4096 → two
        assert launches == {8: 0, 16: 0, 128: 0, 1024: 1, 4: 0, 16384: 2, 65536: 2, 1 << 19: 3}[nb]


def _fold_rounds_of(cnt, keep):
    r = 0
    while cnt > keep:
        cnt = (cnt + 31) // 32; r += 1
    return r, cnt


def test_hot_bucket_folds_keep_their_layout():
    """k_fold_hot_quad + the readers (k_bucket_reduce_quad with keep = 32, k_bucket_reduce<PAIRS> with keep = 1): a bucket's item
    partials stay at its own offset, every round writes the 32:1 sums of the buckets still above `keep` to the other buffer, and each
    bucket finds its partials in buffer (rounds mod 2) with the count fold_rounds_of gives — whatever the other buckets did."""
    rng = random.Random(8)
    for keep in (1, 32):
        counts = [rng.choice([0, 1, 2, 5, 31, 32, 33, 64, 65, 700, 1025, 40000]) for _ in range(40)]
        start = [0]
        for c in counts:
            start.append(start[-1] + c)
        a = [rng.randrange(1 << 30) for _ in range(start[-1])]
        want = [sum(a[start[b]:start[b + 1]]) for b in range(len(counts))]
        bufs = [list(a), [None] * len(a)]
        worst = max(counts)
        rounds = 0
        while worst > keep:                                  # the host loop: rounds for the worst bucket
            src, dst = bufs[rounds & 1], bufs[(rounds + 1) & 1]
            for b, c0 in enumerate(counts):                  # the device: only buckets still above `keep` take part
                cnt, live = c0, True
                for _ in range(rounds):
                    if cnt <= keep:
                        live = False
                    cnt = (cnt + 31) // 32
                if not live or cnt <= keep:
                    continue
                for g in range((cnt + 31) // 32):
                    dst[start[b] + g] = sum(src[start[b] + 32 * g: start[b] + min(32 * g + 32, cnt)])
            worst = (worst + 31) // 32
            rounds += 1
        for b, c0 in enumerate(counts):
            r, cnt = _fold_rounds_of(c0, keep)
            assert r <= rounds and cnt <= keep
            got = sum(bufs[r & 1][start[b]: start[b] + cnt])
            assert got == want[b], (keep, b, c0)

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

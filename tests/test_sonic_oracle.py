"""CPU: oracle/sonic.py (SonicKZG10 commit / batch_open / open_combinations restated) pinned BY DEFINITION on a universal setup
whose trapdoor the test knows — powers_of_beta_g[i] = β^i·G, powers_of_beta_times_gamma_g[i] = γβ^i·G — where every group element
the scheme outputs has a closed form in the scalar field:
    commit(p, r)  = (p(β) + γ·r(β))·G                       (kzg10/mod.rs:98-156)
    commit, bound = (β^{D−d}·p(β) + γβ^{D−d}·r(β))·G        (shifted powers, sonic_pc/data_structures.rs:310-331)
    commit_lagrange(evals) = commit(ifft(evals))            (kzg10/mod.rs:159-206)
    open: w = (q_p(β) + γ·q_r(β))·G, q_f = (f − f(z))/(x − z), random_v = r(z)   (kzg10/mod.rs:220-277)
the last being the verification equation e(C − v·G − γ·v̄·G, H) = e(w, (β − z)·H) in the exponent (kzg10/mod.rs:323-339)."""
import random

import numpy as np

from oracle import bls12_377 as py
from oracle import sonic

R = py.R_MOD
BETA, GAMMA = 0x1234567890ABCDEF1234567890ABCDEF % R, 0xFEDCBA0987654321FEDCBA % R
D = 47                                                   # max_degree of the universal parameters


def _srs(cpu):
    g = np.frombuffer(py.affine_bytes(py.G1_GENERATOR), dtype=np.uint8)
    def pts(scalars):
        out = np.zeros((len(scalars), 104), dtype=np.uint8)
        for i, s in enumerate(scalars):
            proj = cpu.g1_mul(g, sonic._scalars([s])[0])
            aff = py.projective_from_bytes(proj.tobytes())
            out[i] = np.frombuffer(py.affine_bytes(aff), dtype=np.uint8)
        return out
    return pts([pow(BETA, i, R) for i in range(D + 1)]), pts([GAMMA * pow(BETA, i, R) % R for i in range(D + 2)])


def _g(cpu, k):
    g = np.frombuffer(py.affine_bytes(py.G1_GENERATOR), dtype=np.uint8)
    return cpu.g1_mul(g, sonic._scalars([k % R])[0])


def _q(f, z):
    """(f(β) − f(z)) / (β − z)"""
    return (sonic.poly_eval(f, BETA) - sonic.poly_eval(f, z)) * pow(BETA - z, -1, R) % R


def test_division_and_axpy_helpers():
    rnd = random.Random(1)
    p = [rnd.randrange(R) for _ in range(9)]
    z = rnd.randrange(R)
    q = sonic.divide_by_linear(p, z)
    # p = q·(x − z) + p(z)
    back = [0] * len(p)
    for i, c in enumerate(q):
        back[i + 1] = (back[i + 1] + c) % R
        back[i] = (back[i] - z * c) % R
    back[0] = (back[0] + sonic.poly_eval(p, z)) % R
    assert back == p
    assert sonic.divide_by_linear([5], z) == [] and sonic.divide_by_linear([], z) == []
    assert sonic.poly_axpy([1, 2], 3, [1, 1, 1]) == [4, 5, 3]


def test_commit_closed_forms(oracle_cpu):
    rnd = random.Random(2)
    powers, gamma = _srs(oracle_cpu)
    ck = sonic.CommitterKey(powers, gamma, supported_degree=40, supported_lagrange_sizes=[8], supported_hiding_bound=1,
                            enforced_degree_bounds=[10, 30])
    f = [rnd.randrange(R) for _ in range(33)]                       # plain, non-hiding
    h = [rnd.randrange(R) for _ in range(20)]; hr = [rnd.randrange(R) for _ in range(3)]      # plain, hiding_bound = 1
    b10 = [rnd.randrange(R) for _ in range(11)]; b10r = [rnd.randrange(R) for _ in range(3)]  # degree bound 10, hiding
    b30 = [rnd.randrange(R) for _ in range(25)]                     # degree bound 30, degree 24, non-hiding
    ev = [rnd.randrange(R) for _ in range(8)]; evr = [rnd.randrange(R) for _ in range(3)]     # Lagrange, hiding
    comms, rands = sonic.commit(ck, [("f", f, None, None, False), ("h", h, None, 1, False), ("b10", b10, 10, 1, False),
                                     ("b30", b30, 30, None, False), ("ev", ev, None, 1, True)], [None, hr, b10r, None, evr])
    E = sonic.poly_eval
    assert (comms[0] == _g(oracle_cpu, E(f, BETA))).all()
    assert (comms[1] == _g(oracle_cpu, E(h, BETA) + GAMMA * E(hr, BETA))).all()
    s10, s30 = pow(BETA, D - 10, R), pow(BETA, D - 30, R)
    assert (comms[2] == _g(oracle_cpu, s10 * (E(b10, BETA) + GAMMA * E(b10r, BETA)))).all()
    assert (comms[3] == _g(oracle_cpu, s30 * E(b30, BETA))).all()
    coeffs = py.ifft(ev)
    assert (comms[4] == _g(oracle_cpu, E(coeffs, BETA) + GAMMA * E(evr, BETA))).all()
    assert rands == [None, hr, b10r, None, evr]
    # empty polynomial commits to the identity
    assert (sonic.commit(ck, [("z", [], None, None, False)])[0][0] == sonic.INFINITY).all()


def test_open_satisfies_the_kzg_equation_in_the_exponent(oracle_cpu):
    rnd = random.Random(3)
    powers, gamma = _srs(oracle_cpu)
    ck = sonic.CommitterKey(powers, gamma, supported_degree=40, supported_hiding_bound=1)
    p = [rnd.randrange(R) for _ in range(30)]; r = [rnd.randrange(R) for _ in range(3)]
    z = rnd.randrange(R)
    w, rv = sonic.kzg_open(ck.powers_of_beta_g, ck.powers_of_beta_times_gamma_g, p, z, r)
    assert (w == _g(oracle_cpu, _q(p, z) + GAMMA * _q(r, z))).all() and rv == sonic.poly_eval(r, z)
    # C − v·G − γ·v̄·G = (β − z)·w   (kzg10/mod.rs:323-339 without the pairing)
    c = sonic.poly_eval(p, BETA) + GAMMA * sonic.poly_eval(r, BETA)
    lhs = (c - sonic.poly_eval(p, z) - GAMMA * rv) % R
    assert lhs == (BETA - z) * (_q(p, z) + GAMMA * _q(r, z)) % R
    w2, rv2 = sonic.kzg_open(ck.powers_of_beta_g, ck.powers_of_beta_times_gamma_g, p, z)
    assert rv2 is None and (w2 == _g(oracle_cpu, _q(p, z))).all()


def test_batch_open_and_open_combinations(oracle_cpu):
    rnd = random.Random(4)
    powers, gamma = _srs(oracle_cpu)
    ck = sonic.CommitterKey(powers, gamma, supported_degree=40, supported_hiding_bound=1, enforced_degree_bounds=[12])
    P = {name: [rnd.randrange(R) for _ in range(n)] for name, n in (("a", 17), ("b", 33), ("c", 9), ("g", 13))}
    Rz = {"a": [rnd.randrange(R) for _ in range(3)], "b": None, "c": [rnd.randrange(R) for _ in range(3)], "g": [rnd.randrange(R) for _ in range(3)]}
    beta_pt, gamma_pt = rnd.randrange(R), rnd.randrange(R)
    # batch_open: a, c at "beta"; b, c at "gamma" — challenges are consumed per point name (sorted), labels sorted, then one randomizer
    chal = [rnd.randrange(1 << 128) for _ in range(6)]
    qs = [("c", ("beta", beta_pt)), ("a", ("beta", beta_pt)), ("b", ("gamma", gamma_pt)), ("c", ("gamma", gamma_pt))]
    proofs = sonic.batch_open(ck, {k: (P[k], Rz[k]) for k in P}, qs, iter(chal))
    comb0 = sonic.poly_axpy(sonic.poly_axpy([], chal[0], P["a"]), chal[1], P["c"])
    rnd0 = sonic.poly_axpy(sonic.poly_axpy([], chal[0], Rz["a"]), chal[1], Rz["c"])
    comb1 = sonic.poly_axpy(sonic.poly_axpy([], chal[3], P["b"]), chal[4], P["c"])
    rnd1 = sonic.poly_axpy([], chal[4], Rz["c"])
    assert (proofs[0][0] == _g(oracle_cpu, _q(comb0, beta_pt) + GAMMA * _q(rnd0, beta_pt))).all() and proofs[0][1] == sonic.poly_eval(rnd0, beta_pt)
    assert (proofs[1][0] == _g(oracle_cpu, _q(comb1, gamma_pt) + GAMMA * _q(rnd1, gamma_pt))).all() and proofs[1][1] == sonic.poly_eval(rnd1, gamma_pt)
    # open_combinations: lc1 = 3·a − 2·b + 7 (constant term not committed), lc2 = g alone (degree bound 12, coefficient one)
    lcs = [("lc1", [(3, "a"), (R - 2, "b"), (7, None)]), ("lc2", [(1, "g")])]
    polys = {"a": (P["a"], Rz["a"], None), "b": (P["b"], Rz["b"], None), "g": (P["g"], Rz["g"], 12)}
    chal2 = [rnd.randrange(1 << 128) for _ in range(3)]
    out = sonic.open_combinations(ck, lcs, polys, [("lc1", ("beta", beta_pt)), ("lc2", ("beta", beta_pt))], iter(chal2))
    lc1 = sonic.poly_axpy(sonic.poly_axpy([], 3, P["a"]), R - 2, P["b"]); lc1r = sonic.poly_axpy([], 3, Rz["a"])
    comb = sonic.poly_axpy(sonic.poly_axpy([], chal2[0], lc1), chal2[1], P["g"])
    combr = sonic.poly_axpy(sonic.poly_axpy([], chal2[0], lc1r), chal2[1], Rz["g"])
    assert len(out) == 1
    assert (out[0][0] == _g(oracle_cpu, _q(comb, beta_pt) + GAMMA * _q(combr, beta_pt))).all() and out[0][1] == sonic.poly_eval(combr, beta_pt)

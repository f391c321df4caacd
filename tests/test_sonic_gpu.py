"""GPU: snarkvm_b200.sonic_pc (SonicKZG10 commit / batch_open / open_combinations on device-resident operands) against
oracle/sonic.py on the same inputs, and against the closed forms a known trapdoor gives (tests/test_sonic_oracle.py)."""
import random

import numpy as np
import pytest

from oracle import bls12_377 as py
from oracle import sonic as osonic

from helpers import fr_ints_to_mont_array

pytestmark = pytest.mark.gpu
R = py.R_MOD
BETA, GAMMA = 0x1234567890ABCDEF1234567890ABCDEF % R, 0xFEDCBA0987654321FEDCBA % R


def _dev(vals):
    import torch
    return torch.from_numpy(fr_ints_to_mont_array(vals).view(np.int64).copy()).cuda() if len(vals) else torch.zeros((0, 4), dtype=torch.int64, device="cuda")


def _mont_to_int(x):
    return py.fr_from_mont(py.from_limbs(np.asarray(x, dtype=np.uint64).reshape(4)))


@pytest.fixture(scope="module")
def srs():
    from snarkvm_b200 import sonic_pc
    D = 2047
    powers, gamma = sonic_pc.synthetic_srs(D, BETA, GAMMA)
    return D, powers, gamma, powers.cpu().numpy(), gamma.cpu().numpy()


def test_generated_powers_are_the_powers_of_beta(oracle_cpu, srs):
    D, powers, gamma, hp, hg = srs
    g = np.frombuffer(py.affine_bytes(py.G1_GENERATOR), dtype=np.uint8)
    assert hp.shape == (D + 1, 104) and hg.shape == (D + 2, 104)
    for i in (0, 1, 2, 63, 64, 1000, D):
        want = py.projective_from_bytes(oracle_cpu.g1_mul(g, osonic._scalars([pow(BETA, i, R)])[0]).tobytes())
        assert hp[i].tobytes() == py.affine_bytes(want), i
        wantg = py.projective_from_bytes(oracle_cpu.g1_mul(g, osonic._scalars([GAMMA * pow(BETA, i, R) % R])[0]).tobytes())
        assert hg[i].tobytes() == py.affine_bytes(wantg), i


def test_commit_round_vs_oracle(oracle_cpu, srs):
    """one pass over plain powers, two shifts, a Lagrange basis and blinding terms (sonic_pc/mod.rs:177-257)"""
    from snarkvm_b200.sonic_pc import CommitterKey, LabeledPolynomial, SonicKZG10
    D, powers, gamma, hp, hg = srs
    rnd = random.Random(11)
    ck = CommitterKey.trim(powers, gamma, supported_degree=2000, supported_lagrange_sizes=[256], supported_hiding_bound=1,
                           enforced_degree_bounds=[500, 1500])
    ock = osonic.CommitterKey(hp, hg, 2000, [256], 1, [500, 1500])
    assert (ck.lagrange_bases_at_beta_g[256].cpu().numpy() == ock.lagrange_bases_at_beta_g[256]).all()
    spec = [("w", 1800, None, 1, False), ("mask", 1200, None, None, False), ("g_1", 501, 500, 1, False), ("g_a", 1400, 1500, None, False),
            ("ev", 256, None, 1, True), ("ev_short", 200, None, None, True), ("zero", 0, None, None, False), ("tiny", 1, None, None, False)]
    vals = {name: [rnd.randrange(R) for _ in range(n)] for name, n, *_ in spec}
    blind = {name: ([rnd.randrange(R) for _ in range(hb + 2)] if hb is not None else None) for name, _, _, hb, _ in spec}
    polys = [LabeledPolynomial(name, _dev(vals[name]), db, hb, lag) for name, _, db, hb, lag in spec]
    comms, rands = SonicKZG10.commit(ck, polys, [None if blind[n] is None else _dev(blind[n]) for n, *_ in spec])
    want, _ = osonic.commit(ock, [(name, vals[name], db, hb, lag) for name, _, db, hb, lag in spec], [blind[n] for n, *_ in spec])
    for i, (name, *_rest) in enumerate(spec):
        assert (comms[i] == want[i]).all(), name
    assert [r.is_hiding() for r in rands] == [b is not None for b in (blind[n] for n, *_ in spec)]
    # closed form of the degree-bounded hiding commitment
    E = osonic.poly_eval
    g = np.frombuffer(py.affine_bytes(py.G1_GENERATOR), dtype=np.uint8)
    k = pow(BETA, D - 500, R) * (E(vals["g_1"], BETA) + GAMMA * E(blind["g_1"], BETA)) % R
    assert (comms[2] == oracle_cpu.g1_mul(g, osonic._scalars([k])[0])).all()
    # errors the reference raises
    with pytest.raises(ValueError):
        SonicKZG10.commit(ck, [LabeledPolynomial("bad", _dev(vals["w"]), 700, None)])          # bound not enforced
    with pytest.raises(ValueError):
        SonicKZG10.commit(ck, [LabeledPolynomial("bad", _dev(vals["w"]), 500, None)])          # degree above its bound
    with pytest.raises(ValueError):
        SonicKZG10.commit(ck, [LabeledPolynomial("bad", _dev(vals["w"]), None, 1)])            # hiding without a blinding polynomial


def test_batch_open_and_open_combinations_vs_oracle(oracle_cpu, srs):
    from snarkvm_b200.sonic_pc import CommitterKey, LabeledPolynomial, Randomness, SonicKZG10
    D, powers, gamma, hp, hg = srs
    rnd = random.Random(12)
    ck = CommitterKey.trim(powers, gamma, supported_degree=2000, supported_hiding_bound=1, enforced_degree_bounds=[700])
    ock = osonic.CommitterKey(hp, hg, 2000, (), 1, [700])
    sizes = {"a": 1700, "b": 1999, "c": 900, "g": 640, "h": 1}
    P = {k: [rnd.randrange(R) for _ in range(n)] for k, n in sizes.items()}
    B = {"a": [rnd.randrange(R) for _ in range(3)], "b": None, "c": [rnd.randrange(R) for _ in range(3)], "g": [rnd.randrange(R) for _ in range(3)], "h": None}
    bounds = {"a": None, "b": None, "c": None, "g": 700, "h": None}
    labeled = [LabeledPolynomial(k, _dev(P[k]), bounds[k], 1 if B[k] else None) for k in P]
    rands = [Randomness(_dev(B[k]) if B[k] else None) for k in P]
    beta_pt, gamma_pt = rnd.randrange(R), rnd.randrange(R)
    qs = [("c", ("beta", beta_pt)), ("a", ("beta", beta_pt)), ("b", ("gamma", gamma_pt)), ("c", ("gamma", gamma_pt)), ("h", ("gamma", gamma_pt))]
    chal = [rnd.randrange(1 << 128) for _ in range(7)]
    got = SonicKZG10.batch_open(ck, labeled, qs, rands, iter(chal))
    want = osonic.batch_open(ock, {k: (P[k], B[k]) for k in P}, qs, iter(chal))
    assert len(got) == len(want) == 2
    for (gw, gv), (ww, wv) in zip(got, want):
        assert (gw == ww).all()
        assert (gv is None) == (wv is None) and (gv is None or _mont_to_int(gv) == wv)
    # the first proof against the closed form: w = (q_p(β) + γ·q_r(β))·G for the combined polynomial
    comb = osonic.poly_axpy(osonic.poly_axpy([], chal[0], P["a"]), chal[1], P["c"]); combr = osonic.poly_axpy(osonic.poly_axpy([], chal[0], B["a"]), chal[1], B["c"])
    q = lambda f, z: (osonic.poly_eval(f, BETA) - osonic.poly_eval(f, z)) * pow(BETA - z, -1, R) % R
    g = np.frombuffer(py.affine_bytes(py.G1_GENERATOR), dtype=np.uint8)
    assert (got[0][0] == oracle_cpu.g1_mul(g, osonic._scalars([(q(comb, beta_pt) + GAMMA * q(combr, beta_pt)) % R])[0])).all()
    # linear combinations (the Varuna prover's last step, varuna.rs:586-594)
    lcs = [("lc1", [(3, "a"), (R - 2, "b"), (7, None)]), ("lc2", [(1, "g")]), ("lc3", [(5, "c"), (1, "h")])]
    qs2 = [("lc1", ("beta", beta_pt)), ("lc2", ("beta", beta_pt)), ("lc3", ("gamma", gamma_pt)), ("lc1", ("gamma", gamma_pt))]
    chal2 = [rnd.randrange(1 << 128) for _ in range(6)]
    got2 = SonicKZG10.open_combinations(ck, lcs, labeled, rands, qs2, iter(chal2))
    want2 = osonic.open_combinations(ock, lcs, {k: (P[k], B[k], bounds[k]) for k in P}, qs2, iter(chal2))
    assert len(got2) == len(want2) == 2
    for (gw, gv), (ww, wv) in zip(got2, want2):
        assert (gw == ww).all()
        assert (gv is None) == (wv is None) and (gv is None or _mont_to_int(gv) == wv)
    with pytest.raises(ValueError):
        SonicKZG10.open_combinations(ck, [("bad", [(1, "g"), (1, "a")])], labeled, rands, [("bad", ("beta", beta_pt))], iter(chal2))   # EquationHasDegreeBounds


def test_commit_closed_form_at_2_16(oracle_cpu):
    """a 2^16-coefficient polynomial with degree bound and hiding against a generated SRS: (β^{D−d}·(p(β) + γ·r(β)))·G"""
    from snarkvm_b200.sonic_pc import CommitterKey, LabeledPolynomial, SonicKZG10, synthetic_srs
    from helpers import random_canonical_fr
    import torch
    D = (1 << 16) + 99
    powers, gamma = synthetic_srs(D, BETA, GAMMA)
    ck = CommitterKey.trim(powers, gamma, supported_degree=D, supported_hiding_bound=1, enforced_degree_bounds=[(1 << 16) - 1])
    coeffs = random_canonical_fr(1 << 16, seed=5)                   # Montgomery images
    ints = [py.fr_from_mont(py.from_limbs(r)) for r in coeffs]
    rb = [3, 5, 7]
    p = LabeledPolynomial("p", torch.from_numpy(coeffs.view(np.int64)).cuda(), (1 << 16) - 1, 1)
    q = LabeledPolynomial("q", torch.from_numpy(coeffs.view(np.int64)).cuda(), None, None)
    comms, _ = SonicKZG10.commit(ck, [p, q], [_dev(rb), None])
    g = np.frombuffer(py.affine_bytes(py.G1_GENERATOR), dtype=np.uint8)
    pv = osonic.poly_eval(ints, BETA)
    k = pow(BETA, D - ((1 << 16) - 1), R) * (pv + GAMMA * osonic.poly_eval(rb, BETA)) % R
    assert (comms[0] == oracle_cpu.g1_mul(g, osonic._scalars([k])[0])).all()
    assert (comms[1] == oracle_cpu.g1_mul(g, osonic._scalars([pv])[0])).all()

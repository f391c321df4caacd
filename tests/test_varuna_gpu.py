"""GPU parity for SURVEY §8 f3: the device-resident Varuna prover rounds (snarkvm_b200/varuna.py) against
(1) the reference's own golden vectors for circuit_0 (resources/circuit_0/polynomials/*.txt, snark/varuna/tests.rs:623-803) and
(2) the Python restatement of the rounds (oracle/varuna.py, itself pinned to those vectors) on larger TestCircuits, batches of
instances and random sparse R1CS matrices with general coefficients."""
import random

import numpy as np
import pytest

from oracle import varuna as ov

pytestmark = pytest.mark.gpu
ROUND_NAMES = ("w", "z", "h_0", "g_1", "h_1", "g_a", "g_b", "g_c", "h_2")


def _device_run(circuit, zs, ch, mask=None):
    from snarkvm_b200 import varuna as dv
    alpha, eta_b, eta_c, beta, deltas = ch
    p = dv.Prover(circuit, zs)
    if mask is not None:
        p.set_mask_poly(*mask)
    p.first_round(); p.assignments(); p.second_round()
    p.third_round(alpha, eta_b, eta_c)
    p.fourth_round(alpha, beta)
    p.fifth_round(deltas)
    out = {"w": [dv.trimmed(x) for x in p.w_polys], "z": [dv.trimmed(x) for x in p.z_polys], "h_0": dv.trimmed(p.h_0),
           "g_1": dv.trimmed(p.g_1), "h_1": dv.trimmed(p.h_1), "g_a": dv.trimmed(p.gs[0]), "g_b": dv.trimmed(p.gs[1]),
           "g_c": dv.trimmed(p.gs[2]), "h_2": dv.trimmed(p.h_2), "third_sums": p.third_sums, "fourth_sums": p.fourth_sums}
    return out


def _oracle_run(circuit, instances, ch, mask=None):
    alpha, eta_b, eta_c, beta, deltas = ch
    p = ov.Prover(circuit, instances)
    if mask is not None:
        p.set_mask_poly(*mask)
    p.first_round(); p.assignments(); p.second_round()
    p.third_round(alpha, eta_b, eta_c)
    p.fourth_round(alpha, beta)
    p.fifth_round(deltas)
    return {"w": p.w_polys, "z": p.z_polys, "h_0": p.h_0, "g_1": p.g_1, "h_1": p.h_1, "g_a": p.gs[0], "g_b": p.gs[1], "g_c": p.gs[2],
            "h_2": p.h_2, "third_sums": p.third_sums, "fourth_sums": p.fourth_sums}


def test_circuit_0_against_reference_vectors(golden):
    """the nine prover polynomials of the reference's test_varuna_with_prover_test_vectors, computed on the device"""
    from snarkvm_b200 import varuna as dv
    kat = golden["varuna_circuit_0_prover"]
    a, b = kat["witness_a_b"]
    circuit, z = dv.test_circuit_csr(a, b, 3, 7, 7, "cuda")
    ch = [int(x) for x in kat["challenges"]]
    got = _device_run(circuit, [z], (ch[0], ch[2], ch[3], ch[4], ch[5:8]))
    ints = lambda v: [int(x) for x in v]        # noqa: E731
    assert got["w"][0] == ints(kat["w_lde"])
    assert got["z"][0] == ints(kat["z_lde"])
    for name in ("h_0", "g_1", "h_1", "g_a", "g_b", "h_2"):
        assert got[name] == ints(kat[name]), name
    assert got["g_b"] == ints(kat["g_c"])           # tests.rs:736 writes g_b into g_c.txt


def _challenges(rng):
    r = lambda: rng.randrange(2, ov.R)          # noqa: E731
    return r(), r(), r(), r(), [r(), r(), r()]


@pytest.mark.parametrize("mul_depth,num_constraints,num_variables,batch", [(1, 16, 16, 1), (3, 100, 70, 2), (2, 1 << 10, (1 << 10) - 10, 3), (5, 3000, 1 << 12, 1)])
def test_test_circuit_rounds_vs_oracle(mul_depth, num_constraints, num_variables, batch):
    """TestCircuit at several shapes (constraint domain ≠ variable domain ≠ non-zero domain), batches of instances with different
    witnesses: every round polynomial and every sumcheck claim equals the oracle's"""
    from snarkvm_b200 import varuna as dv
    rng = random.Random(num_constraints * 31 + batch)
    wit = [(rng.randrange(2, ov.R), rng.randrange(2, ov.R)) for _ in range(batch)]
    o_circuit = ov.Circuit(ov.test_circuit(wit[0][0], wit[0][1], mul_depth, num_constraints, num_variables))
    o_inst = [ov.test_circuit(a, b, mul_depth, num_constraints, num_variables) for a, b in wit]
    zs, circuit = [], None
    for a, b in wit:
        circuit, z = dv.test_circuit_csr(a, b, mul_depth, num_constraints, num_variables, "cuda")
        zs.append(z)
    ch = _challenges(rng)
    got, want = _device_run(circuit, zs, ch), _oracle_run(o_circuit, o_inst, ch)
    for k in want:
        assert got[k] == want[k], k


def test_random_sparse_r1cs_vs_oracle():
    """general CSR matrices: rows with 1–4 entries and arbitrary coefficients, repeated columns merged by the indexer; C is chosen so
    the instance is satisfied (the rowcheck remainder must vanish)"""
    from snarkvm_b200 import varuna as dv
    rng = random.Random(77)
    n_pub, n_prv, n_con = 4, 53, 90
    cs = ov.ConstraintSystem()
    for _ in range(n_pub - 1):
        cs.alloc_input(rng.randrange(ov.R))
    for _ in range(n_prv):
        cs.alloc(rng.randrange(1, ov.R))
    var = lambda: ("pub", rng.randrange(n_pub)) if rng.random() < 0.2 else ("priv", rng.randrange(n_prv))     # noqa: E731
    val = lambda v: cs.public[v[1]] if v[0] == "pub" else cs.private[v[1]]                                      # noqa: E731
    for _ in range(n_con):
        la = [(rng.randrange(1, ov.R), var()) for _ in range(rng.randrange(1, 5))]
        lb = [(rng.randrange(1, ov.R), var()) for _ in range(rng.randrange(1, 4))]
        az = sum(c * val(v) for c, v in la) % ov.R
        bz = sum(c * val(v) for c, v in lb) % ov.R
        j = ("priv", rng.randrange(n_prv))
        lc = [(az * bz % ov.R * pow(val(j), -1, ov.R) % ov.R, j)]
        cs.enforce(la, lb, lc)
    import copy
    o_circuit = ov.Circuit(copy.deepcopy(cs))
    mats = []
    for m in (o_circuit.a, o_circuit.b, o_circuit.c):
        row_ptr = np.concatenate([[0], np.cumsum([len(r) for r in m])])
        cols = np.array([c for r in m for _, c in r], dtype=np.int64)
        vals = np.array([dv._mont(v) for r in m for v, _ in r], dtype=np.uint64).reshape(-1, 4)
        mats.append(dv.Matrix(row_ptr, cols, vals, "cuda"))
    circuit = dv.Circuit(mats[0], mats[1], mats[2], o_circuit.num_public, o_circuit.num_variables)
    import torch
    z = torch.from_numpy(np.array([dv._mont(v) for v in cs.public + cs.private], dtype=np.uint64).reshape(-1, 4).view(np.int64)).cuda()
    ch = _challenges(rng)
    got, want = _device_run(circuit, [z], ch), _oracle_run(o_circuit, [cs], ch)
    for k in want:
        assert got[k] == want[k], k


def test_hiding_mode_mask_polynomial_vs_oracle():
    """VarunaHidingMode's mask polynomial (first.rs:102-127) enters h_1 and g_1 in the third round (third.rs:207-213)"""
    from snarkvm_b200 import varuna as dv
    rng = random.Random(5)
    a, b = rng.randrange(2, ov.R), rng.randrange(2, ov.R)
    o_circuit = ov.Circuit(ov.test_circuit(a, b, 2, 200, 180))
    circuit, z = dv.test_circuit_csr(a, b, 2, 200, 180, "cuda")
    mask = ([rng.randrange(ov.R) for _ in range(4)], [rng.randrange(ov.R) for _ in range(6)])
    ch = _challenges(rng)
    got, want = _device_run(circuit, [z], ch, mask), _oracle_run(o_circuit, [ov.test_circuit(a, b, 2, 200, 180)], ch, mask)
    for k in want:
        assert got[k] == want[k], k
    plain = _oracle_run(o_circuit, [ov.test_circuit(a, b, 2, 200, 180)], ch)
    assert want["h_1"] != plain["h_1"] and want["g_1"] != plain["g_1"] and want["h_0"] == plain["h_0"]


def test_linear_combinations_and_openings_vs_oracle(oracle_cpu):
    """the last step of prove_batch (varuna.rs:509-594): construct_linear_combinations on the device prover equals the oracle's
    (whose three checks vanish at α, β, γ: tests/test_varuna_golden.py), and SonicKZG10 commits every oracle with the reference's
    bounds (hiding mode: w, g_1, g_M hide; g_1, g_M are degree-bounded) and opens the combinations — against oracle/sonic.py."""
    import torch
    from oracle import sonic as osonic
    from snarkvm_b200 import varuna as dv
    from snarkvm_b200.sonic_pc import CommitterKey, LabeledPolynomial, SonicKZG10, synthetic_srs
    rng = random.Random(21)
    wit = [(rng.randrange(2, ov.R), rng.randrange(2, ov.R)) for _ in range(2)]
    shape = (3, 300, 200)
    o_circuit = ov.Circuit(ov.test_circuit(wit[0][0], wit[0][1], *shape))
    op = ov.Prover(o_circuit, [ov.test_circuit(a, b, *shape) for a, b in wit])
    zs, circuit = [], None
    for a, b in wit:
        circuit, z = dv.test_circuit_csr(a, b, *shape, "cuda")
        zs.append(z)
    dp = dv.Prover(circuit, zs)
    mask = ([rng.randrange(ov.R) for _ in range(4)], [rng.randrange(ov.R) for _ in range(6)])
    r = lambda: rng.randrange(2, ov.R)          # noqa: E731
    alpha, eta_b, eta_c, beta, gamma, deltas, combs = r(), r(), r(), r(), r(), [r(), r(), r()], [1, r()]
    for p in (op, dp):
        p.set_mask_poly(*mask)
        p.first_round(); p.assignments(); p.second_round(1, combs)
        p.third_round(alpha, eta_b, eta_c, 1, combs)
        p.fourth_round(alpha, beta)
        p.fifth_round(deltas)
    want_lcs, want_qs = op.linear_combinations(alpha, eta_b, eta_c, beta, deltas, gamma, 1, combs)
    got_lcs, got_qs = dp.linear_combinations(alpha, eta_b, eta_c, beta, deltas, gamma, 1, combs)
    assert got_lcs == want_lcs and got_qs == want_qs
    opolys, dpolys = op.polynomials(), dp.polynomials()
    assert sorted(opolys) == sorted(dpolys)
    for k in opolys:
        assert dv.trimmed(dpolys[k]) == opolys[k], k
    # commitments and openings over an SRS with a known trapdoor
    D = 2047
    BETA, GAMMA = 0x1234567890ABCDEF % ov.R, 0xFEDCBA09 % ov.R
    powers, gpowers = synthetic_srs(D, BETA, GAMMA)
    V, K = o_circuit.variable_domain, o_circuit.max_non_zero_domain
    bound_g1, bound_gm = V.size - 2, K.size - 2                                  # third.rs / fourth.rs polynomial infos
    bounds = {"g_1": bound_g1, "g_a": o_circuit.non_zero_domains[0].size - 2, "g_b": o_circuit.non_zero_domains[1].size - 2,
              "g_c": o_circuit.non_zero_domains[2].size - 2}
    hiding = {"w_0", "w_1", "g_1", "g_a", "g_b", "g_c"}
    ck = CommitterKey.trim(powers, gpowers, supported_degree=D, supported_hiding_bound=1, enforced_degree_bounds=sorted(set(bounds.values())))
    ock = osonic.CommitterKey(powers.cpu().numpy(), gpowers.cpu().numpy(), D, (), 1, sorted(set(bounds.values())))
    labels = sorted(opolys)
    blind = {k: ([rng.randrange(ov.R) for _ in range(3)] if k in hiding else None) for k in labels}
    to_dev = lambda v: torch.from_numpy(np.array([dv._mont(x) for x in v], dtype=np.uint64).reshape(-1, 4).view(np.int64)).cuda()   # noqa: E731
    def fit(k):                                                                  # device polynomials may carry trailing zeros: cut to the bound
        t = dpolys[k]
        return t[: bounds[k] + 1].contiguous() if k in bounds and t.shape[0] > bounds[k] + 1 else t.contiguous()
    labeled = [LabeledPolynomial(k, fit(k), bounds.get(k), 1 if k in hiding else None) for k in labels]
    comms, rands = SonicKZG10.commit(ck, labeled, [None if blind[k] is None else to_dev(blind[k]) for k in labels])
    want_comms, _ = osonic.commit(ock, [(k, opolys[k], bounds.get(k), 1 if k in hiding else None, False) for k in labels], [blind[k] for k in labels])
    for i, k in enumerate(labels):
        assert (comms[i] == want_comms[i]).all(), k
    chal = [rng.randrange(1 << 128) for _ in range(len(got_lcs) + 3)]
    got = SonicKZG10.open_combinations(ck, got_lcs, labeled, rands, got_qs, iter(chal))
    want = osonic.open_combinations(ock, want_lcs, {k: (opolys[k], blind[k], bounds.get(k)) for k in labels}, want_qs, iter(chal))
    assert len(got) == len(want) == 3
    for (gw, gv), (ww, wv) in zip(got, want):
        assert (gw == ww).all()
        assert (gv is None) == (wv is None)
        if gv is not None:
            from oracle import bls12_377 as py
            assert py.fr_from_mont(py.from_limbs(np.asarray(gv, dtype=np.uint64))) == wv

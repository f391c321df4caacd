"""Pins the oracle (oracle/bls12_377.py and oracle/oracle.c) against the reference's own constants,
known-answer data and real SRS points (tests/golden/, extracted by tests/golden/make_golden.py)."""
import os
import random

import numpy as np
import pytest

from oracle import bls12_377 as py

from helpers import affine_array, fr_ints_to_mont_array, mont_array_to_fr_ints, scalars_from_ints

HERE = os.path.dirname(os.path.abspath(__file__))


def test_fr_constants(golden, oracle_cpu):
    g = golden["fr"]
    assert py.from_limbs(g["MODULUS"]) == py.R_MOD
    assert py.from_limbs(g["R"]) == py.FR_MONT_R
    assert py.from_limbs(g["R2"]) == py.FR_MONT_R2
    assert g["INV"] == py.FR_INV64
    assert g["TWO_ADICITY"] == py.FR_TWO_ADICITY
    assert py.fr_from_mont(py.from_limbs(g["GENERATOR"])) == py.FR_GENERATOR
    assert py.fr_from_mont(py.from_limbs(g["TWO_ADIC_ROOT_OF_UNITY"])) == py.FR_TWO_ADIC_ROOT
    # (r - 1) = 2^47 · T and GENERATOR^T is the 2-adic root (how the constant is defined)
    T = py.from_limbs(g["T"])
    assert (py.R_MOD - 1) == T << 47
    assert pow(py.FR_GENERATOR, T, py.R_MOD) == py.FR_TWO_ADIC_ROOT
    # C oracle agrees on Montgomery one: to_mont(1) == R
    one = oracle_cpu.fr_to_mont(np.array([[1, 0, 0, 0]], dtype=np.uint64))
    assert py.from_limbs(one[0]) == py.FR_MONT_R


def test_fr_powers_of_roots_of_unity_table(golden, oracle_cpu):
    """POWERS_OF_ROOTS_OF_UNITY[k] = root^(2^k) in Montgomery form (fr.rs:60-107; test_powers_of_root_of_unity
    fr.rs:212-239): pins Fr Montgomery multiplication of the C oracle on 46 reference-made values."""
    tab = golden["fr"]["POWERS_OF_ROOTS_OF_UNITY"]
    assert len(tab) == 46
    cur = np.array(tab[0], dtype=np.uint64)
    assert py.fr_from_mont(py.from_limbs(tab[0])) == py.FR_TWO_ADIC_ROOT
    for k in range(1, len(tab)):
        cur = oracle_cpu.fr_mul(cur, cur)
        assert [int(x) for x in cur] == tab[k], k
    last = oracle_cpu.fr_mul(cur, cur)      # root^(2^46) = -1
    assert py.fr_from_mont(py.from_limbs(last)) == py.R_MOD - 1


def test_fq_constants_and_root_kat(golden, oracle_cpu):
    """curves/src/bls12_377/tests.rs:460-476: multiplicative_generator^T == two_adic_root_of_unity, computed
    here with the C oracle's Fq Montgomery multiplication (square-and-multiply over the reference's T)."""
    g = golden["fq"]
    assert py.from_limbs(g["MODULUS"]) == py.Q_MOD
    assert py.from_limbs(g["R"]) == py.FQ_MONT_R
    assert py.from_limbs(g["R2"]) == py.FQ_MONT_R2
    assert g["INV"] == py.FQ_INV64
    T = py.from_limbs(g["T"])
    assert (py.Q_MOD - 1) == T << g["TWO_ADICITY"]
    base = np.array(g["GENERATOR"], dtype=np.uint64)
    acc = np.array(g["R"], dtype=np.uint64)
    for bit in bin(T)[2:]:
        acc = oracle_cpu.fq_mul(acc, acc)
        if bit == "1":
            acc = oracle_cpu.fq_mul(acc, base)
    assert [int(x) for x in acc] == g["TWO_ADIC_ROOT_OF_UNITY"]
    tab = g["POWERS_OF_ROOTS_OF_UNITY"]
    cur = np.array(tab[0], dtype=np.uint64)
    for k in range(1, len(tab)):
        cur = oracle_cpu.fq_mul(cur, cur)
        assert [int(x) for x in cur] == tab[k], k


def test_varuna_domain_kat(golden, oracle_cpu):
    """circuit_0/domain/{R,C,K}.txt hold EvaluationDomain::elements() = ω^i (decimal).  The forward FFT of the
    delta polynomial X (coefficients 0,1,0,…) evaluates to exactly those elements — pins get_root_of_unity and
    the FFT of both oracles."""
    dom = golden["varuna_circuit_0_domain"]
    for name in ("R", "C", "K"):
        elems = [int(s) for s in dom[name]]
        n = len(elems)
        assert n & (n - 1) == 0
        w = py.fr_root_of_unity(n) if n > 1 else 1
        assert elems == [pow(w, i, py.R_MOD) for i in range(n)]
        if n >= 2:
            delta = [0, 1] + [0] * (n - 2)
            assert py.fft(delta) == elems
            got = oracle_cpu.ntt(fr_ints_to_mont_array(delta), 0, 0)
            assert mont_array_to_fr_ints(got) == elems


def test_g1_generator(golden, oracle_cpu):
    g = golden["g1"]
    assert int(g["GENERATOR_X_DEC"]) == py.G1_GEN_X and int(g["GENERATOR_Y_DEC"]) == py.G1_GEN_Y
    assert py.fq_from_mont(py.from_limbs(g["GENERATOR_X_MONT"])) == py.G1_GEN_X
    assert py.fq_from_mont(py.from_limbs(g["GENERATOR_Y_MONT"])) == py.G1_GEN_Y
    assert py.g1_is_on_curve(py.G1_GENERATOR)
    assert oracle_cpu.g1_is_on_curve(affine_array([py.G1_GENERATOR])[0])
    # r·G = ∞ through the C oracle's mul_bits (r itself is a valid 4-limb input)
    out = oracle_cpu.g1_mul(affine_array([py.G1_GENERATOR])[0], scalars_from_ints([py.R_MOD])[0])
    assert out.tobytes() == py.projective_bytes_normalised(None)


def test_g2_constants_and_generator(golden):
    """The G2 oracle's numbers are the reference's (g2.rs, fq2.rs): NONRESIDUE = −5, B' = (0, b1), the generator's four Fq
    coordinates converted out of Montgomery form; the generator is on the curve and has order r (bls12_377/tests.rs:673-678);
    the cofactor times r is the curve order implied by … at least #E'(Fq2) = cofactor·r kills a random curve point."""
    from oracle import g2
    g = golden["g2"]
    assert py.fq_from_mont(py.from_limbs(g["NONRESIDUE_MONT"])) == g2.NONRESIDUE == py.Q_MOD - 5
    b0, b1 = (py.fq_from_mont(py.from_limbs(l)) for l in g["WEIERSTRASS_B_MONT"])
    assert (b0, b1) == g2.G2_B
    gen = ((py.fq_from_mont(py.from_limbs(g["GENERATOR_X_C0_MONT"])), py.fq_from_mont(py.from_limbs(g["GENERATOR_X_C1_MONT"]))),
           (py.fq_from_mont(py.from_limbs(g["GENERATOR_Y_C0_MONT"])), py.fq_from_mont(py.from_limbs(g["GENERATOR_Y_C1_MONT"]))))
    assert gen == g2.G2_GEN
    assert g2.g2_is_on_curve(gen)
    assert g2.g2_mul(gen, py.R_MOD - 1) == g2.g2_neg(gen) and g2.g2_add(g2.g2_mul(gen, py.R_MOD - 1), gen) is None
    # Fq2 arithmetic: u² = −5, inverses
    assert g2.f2_mul((0, 1), (0, 1)) == (py.Q_MOD - 5, 0)
    x = (123456789, 987654321)
    assert g2.f2_mul(x, g2.f2_inv(x)) == (1, 0)
    # layouts round-trip and have the reference sizes
    img = g2.g2_affine_bytes(gen)
    assert len(img) == 200 and g2.g2_affine_from_bytes(img) == gen and len(g2.g2_projective_bytes_normalised(gen)) == 288
    assert g2.g2_affine_from_bytes(g2.g2_affine_bytes(None)) is None


def test_g2_standard_msm_matches_naive_sum():
    """standard::msm restated (standard.rs:24-118) against Σ s_i·P_i by double-and-add (msm_naive, mod.rs:52-57) — the check the
    reference's own test makes (mod.rs:90-119) — incl. scalars 0, 1 (the unit-scalar shortcut), r − 1, repeated and opposite points"""
    import random
    from oracle import g2
    rnd = random.Random(11)
    for n in (1, 7, 40):
        ks = [rnd.randrange(1, 1 << 30) for _ in range(n)]
        bases = [g2.g2_mul(g2.G2_GEN, k) for k in ks]
        sc = [rnd.randrange(py.R_MOD) for _ in range(n)]
        if n >= 7:
            sc[0], sc[1], sc[2] = 0, 1, py.R_MOD - 1
            bases[3] = bases[4]; ks[3] = ks[4]
            bases[5] = g2.g2_neg(bases[6]); ks[5] = -ks[6]; sc[5] = sc[6]
        want = g2.g2_mul(g2.G2_GEN, sum(k * s for k, s in zip(ks, sc)) % py.R_MOD)
        assert g2.standard_msm(bases, sc) == want
        assert g2.msm_naive(bases, sc) == want


def _srs_points(count):
    with open(os.path.join(HERE, "golden", "powers_of_beta_15_first512.usrs"), "rb") as f:
        return py.parse_usrs_points(f.read(), count)


def test_real_srs_points(oracle_cpu):
    """First points of the mainnet powers-of-beta-15.usrs: point 0 is the generator; all are on the curve and in
    the order-r subgroup (r·P = ∞) according to both oracles."""
    pts = _srs_points(512)
    assert pts[0] == py.G1_GENERATOR
    assert all(py.g1_is_on_curve(p) for p in pts)
    arr = affine_array(pts)
    assert all(oracle_cpu.g1_is_on_curve(arr[i]) for i in range(0, 512, 7))
    r = scalars_from_ints([py.R_MOD])[0]
    for i in (1, 2, 17, 511):
        assert oracle_cpu.g1_mul(arr[i], r).tobytes() == py.projective_bytes_normalised(None)
        assert py.g1_mul(pts[i], py.R_MOD) is None


def test_msm_on_real_srs_all_algorithms(oracle_cpu):
    """A KZG-style commitment over real powers: batched (the G1 path of VariableBase::msm), standard and naive
    restatements agree with each other and with the Python root of trust."""
    rng = random.Random(2024)
    pts = _srs_points(256)
    sc = [rng.randrange(py.R_MOD) for _ in range(256)]
    sc[0], sc[1], sc[2] = 0, 1, py.R_MOD - 1
    bases, scal = affine_array(pts), scalars_from_ints(sc)
    expect = py.projective_bytes_normalised(py.msm_naive(pts[:64], sc[:64]))
    for algo in (0, 1, 2):
        assert oracle_cpu.msm(bases[:64], scal[:64], algo).tobytes() == expect
    full = [oracle_cpu.msm(bases, scal, algo).tobytes() for algo in (0, 1, 2)]
    assert full[0] == full[1] == full[2]


@pytest.mark.parametrize("n", [1, 2, 5, 14, 15, 31, 32, 50, 100])
def test_msm_oracle_c_vs_python(oracle_cpu, n):
    """test_msm (variable_base/mod.rs:90-107) restated: batched and standard vs msm_naive, plus edge inputs."""
    rng = random.Random(n)
    pts = [py.g1_mul(py.G1_GENERATOR, rng.randrange(1, py.R_MOD)) for _ in range(n)]
    sc = [rng.randrange(py.R_MOD) for _ in range(n)]
    if n >= 15:
        pts[3] = None                      # infinity base
        sc[5], sc[6], sc[7] = 0, 1, py.R_MOD - 1
        pts[9], sc[9] = pts[8], sc[8]      # duplicate base and scalar → doubling inside a bucket
        pts[11], sc[11] = py.g1_neg(pts[10]), sc[10]   # P and −P in the same buckets
    expect = py.projective_bytes_normalised(py.msm_naive(pts, sc))
    bases, scal = affine_array(pts), scalars_from_ints(sc)
    for algo in (0, 1, 2):
        assert oracle_cpu.msm(bases, scal, algo).tobytes() == expect, algo


def test_msm_unequal_lengths(oracle_cpu):
    """variable_base_test_with_bls12_unequal_numbers (msm/tests.rs:53-67): fewer scalars than bases."""
    rng = random.Random(5)
    pts = [py.g1_mul(py.G1_GENERATOR, rng.randrange(1, 1 << 64)) for _ in range(40)]
    sc = [rng.randrange(py.R_MOD) for _ in range(33)]
    expect = py.projective_bytes_normalised(py.msm_naive(pts[:33], sc))
    assert oracle_cpu.msm(affine_array(pts), scalars_from_ints(sc), 0).tobytes() == expect


@pytest.mark.parametrize("lg", range(0, 10))
def test_ntt_oracle_c_vs_python(oracle_cpu, lg):
    """parallel_fft_consistency / test_fft_correctness (fft/tests.rs:119-285) restated for sizes 2^0…2^9."""
    rng = random.Random(lg)
    n = 1 << lg
    a = [rng.randrange(py.R_MOD) for _ in range(n)]
    x = fr_ints_to_mont_array(a)
    if lg <= 6:
        assert py.fft(a) == py.dft_horner(a)
    for d, t, f in ((0, 0, py.fft), (1, 0, py.ifft), (0, 1, py.coset_fft), (1, 1, py.coset_ifft)):
        assert mont_array_to_fr_ints(oracle_cpu.ntt(x, d, t)) == f(a), (lg, d, t)


def test_polymul_oracle(oracle_cpu):
    rng = random.Random(9)
    p1 = [rng.randrange(py.R_MOD) for _ in range(5)]
    p2 = [rng.randrange(py.R_MOD) for _ in range(9)]
    prod = [0] * 16
    for i, a in enumerate(p1):
        for j, b in enumerate(p2):
            prod[i + j] = (prod[i + j] + a * b) % py.R_MOD
    got = oracle_cpu.polymul([fr_ints_to_mont_array(p1), fr_ints_to_mont_array(p2)], [], 4)
    assert mont_array_to_fr_ints(got) == prod
    # one polynomial and one evaluation vector: p1 · (interpolant of e) mod (X^16 − 1)
    e = [rng.randrange(py.R_MOD) for _ in range(16)]
    q = py.ifft(e)
    cyc = [0] * 16
    for i, a in enumerate(p1):
        for j, b in enumerate(q):
            cyc[(i + j) % 16] = (cyc[(i + j) % 16] + a * b) % py.R_MOD
    got = oracle_cpu.polymul([fr_ints_to_mont_array(p1)], [fr_ints_to_mont_array(e)], 4)
    assert mont_array_to_fr_ints(got) == cyc

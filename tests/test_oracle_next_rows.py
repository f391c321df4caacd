"""CPU: the next-row oracles (lagrange basis = group iFFT, batch inversion, division by the vanishing polynomial, evaluation)
pinned against the Python big-int root of trust by their DEFINITIONS, not by another FFT."""
import os
import random

import numpy as np

from oracle import bls12_377 as py

from helpers import affine_array, fr_ints_to_mont_array, mont_array_to_fr_ints

HERE = os.path.dirname(os.path.abspath(__file__))


def _points(n, seed):
    rnd = random.Random(seed)
    return [py.g1_mul(py.G1_GENERATOR, rnd.randrange(1, py.R_MOD)) for _ in range(n)]


def test_g1_ifft_matches_definition(oracle_cpu):
    """L_i = n^{-1} Σ_j ω^{-ij} P_j (kzg10/data_structures.rs:68-72 with the iFFT written out)"""
    for n in (1, 2, 8):
        pts = _points(n, seed=n)
        if n == 8:
            pts[3] = None                                     # an ∞ input
        got = oracle_cpu.g1_ifft(affine_array(pts))
        w = py.fr_root_of_unity(n) if n > 1 else 1
        wi, ni = pow(w, -1, py.R_MOD), pow(n, -1, py.R_MOD)
        for i in range(n):
            acc = None
            for j, p in enumerate(pts):
                acc = py.g1_add(acc, py.g1_mul(p, pow(wi, i * j, py.R_MOD) * ni % py.R_MOD))
            assert got[i].tobytes() == py.affine_bytes(acc), (n, i)


def test_g1_ifft_real_srs_sums_to_first_power(oracle_cpu):
    """Σ_i L_i(β)·G = 1·G: the Lagrange basis of the real powers-of-beta sums to powers[0] (the generator)"""
    blob = open(os.path.join(HERE, "golden", "powers_of_beta_15_first512.usrs"), "rb").read()
    pts = py.parse_usrs_points(blob, 64)
    basis = oracle_cpu.g1_ifft(affine_array(pts))
    acc = None
    for row in basis:
        acc = py.g1_add(acc, py.affine_from_bytes(row.tobytes()))
    assert acc == pts[0] == py.G1_GENERATOR


def test_batch_inversion_and_mul(oracle_cpu):
    rnd = random.Random(5)
    vals = [rnd.randrange(py.R_MOD) for _ in range(300)]
    for i in (0, 7, 8, 299):
        vals[i] = 0                                           # zeros are skipped and stay zero (fields/src/lib.rs:107,121)
    coeff = rnd.randrange(1, py.R_MOD)
    got = mont_array_to_fr_ints(oracle_cpu.fr_batch_inversion_and_mul(fr_ints_to_mont_array(vals), fr_ints_to_mont_array([coeff])[0]))
    want = [0 if v == 0 else coeff * pow(v, -1, py.R_MOD) % py.R_MOD for v in vals]
    assert got == want
    assert oracle_cpu.fr_batch_inversion_and_mul(np.zeros((0, 4), dtype=np.uint64), fr_ints_to_mont_array([1])[0]).shape == (0, 4)


def test_divide_by_vanishing_and_evaluate(oracle_cpu):
    rnd = random.Random(6)
    for m, n in ((5, 8), (8, 8), (9, 8), (40, 8), (64, 16), (33, 32), (100, 4)):
        p = [rnd.randrange(py.R_MOD) for _ in range(m)]
        if m == 40:
            p[-3:] = [0, 0, 0]                                # trailing zero coefficients
        q, r = oracle_cpu.poly_divide_by_vanishing(fr_ints_to_mont_array(p), n)
        qi, ri = mont_array_to_fr_ints(q), mont_array_to_fr_ints(r)
        assert len(qi) == max(m - n, 0) and len(ri) == min(m, n)
        # p == q·(x^n − 1) + r
        back = [0] * m
        for i, c in enumerate(qi):
            back[i + n] = (back[i + n] + c) % py.R_MOD
            back[i] = (back[i] - c) % py.R_MOD
        for i, c in enumerate(ri):
            back[i] = (back[i] + c) % py.R_MOD
        assert back == p, (m, n)
        z = rnd.randrange(py.R_MOD)
        want = sum(c * pow(z, i, py.R_MOD) for i, c in enumerate(p)) % py.R_MOD
        assert mont_array_to_fr_ints(oracle_cpu.poly_evaluate(fr_ints_to_mont_array(p), fr_ints_to_mont_array([z])[0]).reshape(1, 4)) == [want]


def test_divide_by_linear(oracle_cpu):
    """witness polynomial: p(x) = q(x)·(x − z) + p(z)"""
    rnd = random.Random(8)
    for m in (1, 2, 3, 17, 200):
        p = [rnd.randrange(py.R_MOD) for _ in range(m)]
        if m == 17:
            p[-2:] = [0, 0]
        for z in (0, 1, rnd.randrange(py.R_MOD)):
            q = mont_array_to_fr_ints(oracle_cpu.poly_divide_by_linear(fr_ints_to_mont_array(p), fr_ints_to_mont_array([z])[0]))
            assert len(q) == m - 1
            pz = sum(c * pow(z, i, py.R_MOD) for i, c in enumerate(p)) % py.R_MOD
            back = [0] * m
            for i, c in enumerate(q):
                back[i + 1] = (back[i + 1] + c) % py.R_MOD
                back[i] = (back[i] - c * z) % py.R_MOD
            back[0] = (back[0] + pz) % py.R_MOD
            assert back == p, (m, z)


def test_sparse_matvec(oracle_cpu):
    rnd = random.Random(9)
    nrows, npub, nprv = 40, 5, 30
    pub = [rnd.randrange(py.R_MOD) for _ in range(npub)]
    prv = [rnd.randrange(py.R_MOD) for _ in range(nprv)]
    row_ptr, cols, vals = [0], [], []
    for r in range(nrows):
        for _ in range(rnd.randrange(0, 6)):
            cols.append(rnd.randrange(npub + nprv))
            vals.append(1 if rnd.random() < 0.3 else rnd.randrange(py.R_MOD))      # coefficient.is_one() shortcut (mod.rs:186)
        row_ptr.append(len(cols))
    got = mont_array_to_fr_ints(oracle_cpu.sparse_matvec(row_ptr, cols, fr_ints_to_mont_array(vals) if vals else np.zeros((0, 4), np.uint64),
                                                         fr_ints_to_mont_array(pub), fr_ints_to_mont_array(prv)))
    x = pub + prv
    want = [sum(vals[e] * x[cols[e]] for e in range(row_ptr[r], row_ptr[r + 1])) % py.R_MOD for r in range(nrows)]
    assert got == want

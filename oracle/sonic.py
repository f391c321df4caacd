"""TEST INFRASTRUCTURE (checker, never a product path): CPU restatement of the SonicKZG10 polynomial-commitment layer.

Follows /root/reference/algorithms/src/polycommit:
    sonic_pc/mod.rs:62-175    trim                 (slices of the universal parameters)
    sonic_pc/mod.rs:177-257   commit               (powers / shifted powers / Lagrange basis per polynomial, kzg10 commit)
    sonic_pc/mod.rs:259-284   combine_for_open     + combine_polynomials :548-565
    sonic_pc/mod.rs:286-342   batch_open
    sonic_pc/mod.rs:413-475   open_combinations
    kzg10/mod.rs:98-156       commit               (msm(powers, coeffs) + msm(gamma powers, blinding))
    kzg10/mod.rs:220-277      compute_witness_polynomial / open_with_witness_polynomial
Polynomials are lists of canonical Python ints (low degree first), group elements the C oracle's images (bases: uint8[n, 104]
affine, results: normalised projective uint64[18]); every MSM is the C restatement of batched::msm (oracle.c).

Pinning: the reference holds no commitment/opening vectors (its tests draw the SRS and the polynomials from an entropy-seeded
RNG and check with pairings, sonic_pc/mod.rs:715-790), so this restatement is pinned BY DEFINITION on a universal setup whose
trapdoor (β, γ) the test knows: commit(p, r) = (p(β) + γ·r(β))·G, w = (q_p(β) + γ·q_r(β))·G with q_f = (f − f(z))/(x − z), which is
the KZG verification equation e(C − v·G − γ·v̄·G, H) = e(w, βH − zH) written in the exponent (tests/test_sonic_oracle.py)."""
from __future__ import annotations

import numpy as np

from . import bls12_377 as py
from . import cpu

R = py.R_MOD


def _scalars(vals) -> np.ndarray:
    a = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        v %= R
        for j in range(4):
            a[i, j] = (v >> (64 * j)) & (2**64 - 1)
    return a


INFINITY = np.frombuffer(py.projective_bytes_normalised(None), dtype=np.uint64).copy()


def msm(bases: np.ndarray, coeffs: list) -> np.ndarray:
    """VariableBase::msm(bases[..len], to_bigint(coeffs)) → normalised projective image"""
    if len(coeffs) == 0:
        return INFINITY.copy()
    if len(coeffs) > bases.shape[0]:
        raise ValueError("check_degree_is_too_large")
    return cpu.msm(np.ascontiguousarray(bases[: len(coeffs)]), _scalars(coeffs), cpu.BATCHED)


def poly_eval(p: list, z: int) -> int:
    acc = 0
    for c in reversed(p):
        acc = (acc * z + c) % R
    return acc


def poly_axpy(acc: list, coeff: int, p: list) -> list:
    out = list(acc) + [0] * max(0, len(p) - len(acc))
    for i, c in enumerate(p):
        out[i] = (out[i] + coeff * c) % R
    return out


def divide_by_linear(p: list, z: int) -> list:
    """quotient of p / (x − z) (kzg10/mod.rs:220-241: the remainder p(z) is dropped)"""
    if len(p) <= 1:
        return []
    q = [0] * (len(p) - 1)
    carry = 0
    for i in range(len(p) - 1, 0, -1):
        carry = (p[i] + carry * z) % R
        q[i - 1] = carry
    return q


class CommitterKey:
    def __init__(self, pp_powers, pp_gamma_powers, supported_degree, supported_lagrange_sizes=(), supported_hiding_bound=1,
                 enforced_degree_bounds=None):
        """sonic_pc/mod.rs:62-175 on host arrays (pp_gamma_powers dense: γβ^i·G for i ≤ max_degree + 1)"""
        self.max_degree = pp_powers.shape[0] - 1
        self.powers_of_beta_g = pp_powers[: supported_degree + 1]
        self.powers_of_beta_times_gamma_g = pp_gamma_powers[: supported_hiding_bound + 2]
        self.enforced_degree_bounds = None
        self.shifted_powers_of_beta_g = None
        self.shifted_powers_of_beta_times_gamma_g = None
        if enforced_degree_bounds is not None:
            b = sorted(set(enforced_degree_bounds))
            self.enforced_degree_bounds = b
            if b:
                assert b[-1] <= supported_degree
                self.shifted_powers_of_beta_g = pp_powers[self.max_degree - b[-1]:]
                self.shifted_powers_of_beta_times_gamma_g = {
                    d: pp_gamma_powers[self.max_degree - d: min(self.max_degree, self.max_degree - d + supported_hiding_bound) + 2] for d in b}
        self.lagrange_bases_at_beta_g = {s: cpu.g1_ifft(np.ascontiguousarray(pp_powers[:s])) for s in supported_lagrange_sizes}

    def shifted_powers(self, bound):
        return self.shifted_powers_of_beta_g[self.enforced_degree_bounds[-1] - bound:], self.shifted_powers_of_beta_times_gamma_g[bound]


def kzg_commit(powers, gamma_powers, coeffs, blinding=None) -> np.ndarray:
    """kzg10/mod.rs:98-156"""
    c = msm(powers, coeffs)
    if blinding:
        c = cpu.g1_add(c, msm(gamma_powers, blinding))
    return c


def commit(ck: CommitterKey, polynomials: list, blindings: list | None = None):
    """polynomials: [(label, coeffs-or-evaluations, degree_bound, hiding_bound, lagrange)] → ([commitment], [blinding or None])"""
    blindings = blindings or [None] * len(polynomials)
    comms, rands = [], []
    for (label, poly, degree_bound, hiding_bound, lagrange), b in zip(polynomials, blindings):
        if lagrange:
            size = 1 << max(len(poly) - 1, 0).bit_length()
            bases, gamma = ck.lagrange_bases_at_beta_g[size], ck.powers_of_beta_times_gamma_g
        elif degree_bound is not None:
            bases, gamma = ck.shifted_powers(degree_bound)
        else:
            bases, gamma = ck.powers_of_beta_g, ck.powers_of_beta_times_gamma_g
        r = list(b) if hiding_bound is not None else None
        if r is not None:
            assert len(r) == hiding_bound + 2
        comms.append(kzg_commit(bases, gamma, poly, r))
        rands.append(r)
    return comms, rands


def kzg_open(powers, gamma_powers, poly, z, blinding=None):
    """kzg10/mod.rs:303-321 → (w, random_v or None)"""
    w = msm(powers, divide_by_linear(poly, z))
    if blinding:
        w = cpu.g1_add(w, msm(gamma_powers, divide_by_linear(blinding, z)))
        return w, poly_eval(blinding, z)
    return w, None


def combine_for_open(polys_rands, challenges):
    poly, rand = [], None
    for p, r in polys_rands:
        ch = int(next(challenges)) % R
        poly = poly_axpy(poly, ch, p)
        if r:
            rand = poly_axpy(rand or [], ch, r)
    return poly, rand


def batch_open(ck: CommitterKey, labeled: dict, query_set: list, challenges):
    """labeled: label → (coeffs, blinding or None); query_set: [(label, (point_name, point))] → [(w, random_v)] by point name"""
    by_point: dict = {}
    for label, (name, point) in query_set:
        by_point.setdefault(name, (point, set()))[1].add(label)
    proofs = []
    for name in sorted(by_point):
        point, labels = by_point[name]
        poly, rand = combine_for_open([labeled[l] for l in sorted(labels)], challenges)
        next(challenges)                                                          # `_randomizer`
        proofs.append(kzg_open(ck.powers_of_beta_g, ck.powers_of_beta_times_gamma_g, poly, point % R, rand))
    return proofs


def open_combinations(ck: CommitterKey, linear_combinations: list, polys: dict, query_set: list, challenges):
    """polys: label → (coeffs, blinding or None, degree_bound); linear_combinations: [(lc_label, [(coeff, label or None)])]"""
    lcs = {}
    for lc_label, terms in linear_combinations:
        poly, rand = [], None
        for coeff, label in terms:
            if label is None:
                continue
            p, r, bound = polys[label]
            if bound is not None:
                assert len(terms) == 1 and coeff % R == 1
            poly = poly_axpy(poly, coeff % R, p)
            if r:
                rand = poly_axpy(rand or [], coeff % R, r)
        lcs[lc_label] = (poly, rand)
    return batch_open(ck, lcs, query_set, challenges)

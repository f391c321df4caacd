"""SonicKZG10 — the polynomial commitment layer the Varuna prover calls — on device-resident operands.

Mirror of /root/reference/algorithms/src/polycommit/sonic_pc:
    mod.rs:62-175     trim                  CommitterKey.trim
    mod.rs:177-257    commit                SonicKZG10.commit          (ONE msm_core pass for the whole round)
    mod.rs:259-284    combine_for_open      SonicKZG10.combine_for_open
    mod.rs:286-342    batch_open            SonicKZG10.batch_open
    mod.rs:413-475    open_combinations     SonicKZG10.open_combinations
    data_structures.rs:310-341   shifted_powers_of_beta_g / lagrange_basis
and of kzg10/mod.rs:98-156 (commit), :220-277 (open) through algorithms.KZG10.

Polynomials are CUDA tensors [m, 4] int64 of Montgomery Fr coefficients (low degree first; trailing zeros allowed); bases are CUDA
uint8 tensors [n, 104] in the reference's Affine layout; commitments and proofs come back as the normalised projective image
(uint64[18]) like every MSM of this package.  Two things the Rust code draws from generators are ARGUMENTS here, because neither
generator can be reproduced outside Rust: the blinding polynomials (`KZGRandomness::rand`, kzg10/data_structures.rs) and the
Fiat-Shamir challenges (`fs_rng.squeeze_short_nonnative_field_element`, a Poseidon sponge on the host) — the latter as an iterator
consumed in the reference's squeeze order.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import device
from .algorithms import KZG10, EvaluationDomain, _fr_int_to_mont, _R_MOD

STRIDE = device.AFFINE_STRIDE


# ---- small polynomial helpers on Montgomery coefficient tensors ----
def _zeros(n: int, dev) -> torch.Tensor:
    return torch.zeros((n, 4), dtype=torch.int64, device=dev)


def poly_axpy(acc: torch.Tensor | None, coeff: int, p: torch.Tensor) -> torch.Tensor:
    """acc + coeff·p (DensePolynomial `+= (coeff, &poly)`, fft/polynomial/dense.rs), acc = None is the zero polynomial"""
    coeff %= _R_MOD
    if p.shape[0] == 0 or coeff == 0:
        return acc if acc is not None else _zeros(0, p.device)
    term = p if coeff == 1 else device.fr_vec_op(p, _fr_int_to_mont(coeff), device.FR_MUL)
    if acc is None or acc.shape[0] == 0:
        return term.clone() if term is p else term
    if acc.shape[0] < term.shape[0]:
        out = term.clone() if term is p else term
        device.fr_vec_op(out[: acc.shape[0]], acc, device.FR_ADD, out=out[: acc.shape[0]])
        return out
    out = acc.clone()
    device.fr_vec_op(out[: term.shape[0]], term, device.FR_ADD, out=out[: term.shape[0]])
    return out


@dataclass
class LabeledPolynomial:
    """polycommit/data_structures.rs LabeledPolynomial / LabeledPolynomialWithBasis: `polynomial` holds monomial coefficients, or
    — with `lagrange=True` — evaluations over the domain of their count (committed against the Lagrange basis of that size)."""
    label: str
    polynomial: torch.Tensor
    degree_bound: int | None = None
    hiding_bound: int | None = None
    lagrange: bool = False


@dataclass
class Randomness:
    """KZGRandomness: the blinding polynomial of a hiding commitment (hiding_bound + 2 coefficients), None = Randomness::empty()"""
    blinding_polynomial: torch.Tensor | None = None

    def is_hiding(self) -> bool:
        return self.blinding_polynomial is not None and self.blinding_polynomial.shape[0] > 0

    def axpy(self, coeff: int, other: "Randomness") -> "Randomness":
        if not other.is_hiding():
            return self
        return Randomness(poly_axpy(self.blinding_polynomial, coeff, other.blinding_polynomial))


@dataclass
class CommitterKey:
    """sonic_pc/data_structures.rs CommitterKey / CommitterUnionKey over an SRS resident in HBM"""
    powers_of_beta_g: torch.Tensor                                   # [supported_degree + 1, 104]
    powers_of_beta_times_gamma_g: torch.Tensor                       # [supported_hiding_bound + 2, 104]
    lagrange_bases_at_beta_g: dict = field(default_factory=dict)     # size → [size, 104]
    shifted_powers_of_beta_g: torch.Tensor | None = None             # pp.powers[max_degree − highest bound …]
    shifted_powers_of_beta_times_gamma_g: dict | None = None         # bound → gamma powers of that shift
    enforced_degree_bounds: list | None = None
    max_degree: int = 0

    @classmethod
    def trim(cls, pp_powers_of_beta_g: torch.Tensor, pp_powers_of_beta_times_gamma_g: torch.Tensor, supported_degree: int,
             supported_lagrange_sizes=(), supported_hiding_bound: int = 1, enforced_degree_bounds=None) -> "CommitterKey":
        """mod.rs:62-175.  `pp_powers_of_beta_times_gamma_g` is dense here (γβ^i·G for every i the reference keeps sparsely)."""
        max_degree = pp_powers_of_beta_g.shape[0] - 1
        bounds = None
        shifted = shifted_gamma = None
        if enforced_degree_bounds is not None:
            bounds = sorted(set(int(b) for b in enforced_degree_bounds))
            if bounds:
                highest = bounds[-1]
                if highest > supported_degree:
                    raise ValueError(f"The highest enforced degree bound {highest} is larger than the supported degree {supported_degree}")
                shifted = pp_powers_of_beta_g[max_degree - highest:]
                shifted_gamma = {}
                for b in bounds:
                    shift = max_degree - b
                    shifted_gamma[b] = pp_powers_of_beta_times_gamma_g[shift: min(max_degree, shift + supported_hiding_bound) + 2]
        gamma = pp_powers_of_beta_times_gamma_g[: supported_hiding_bound + 2]
        if gamma.shape[0] != supported_hiding_bound + 2:
            raise ValueError("HidingBoundToolarge")
        bases = {}
        for size in supported_lagrange_sizes:
            if size & (size - 1):
                raise ValueError(f"The Lagrange basis size ({size}) is not a power of two")
            if size > max_degree + 1:
                raise ValueError(f"The Lagrange basis size ({size}) is larger than the supported degree ({max_degree + 1})")
            bases[size] = device.lagrange_basis(pp_powers_of_beta_g[:size].contiguous())
        return cls(pp_powers_of_beta_g[: supported_degree + 1], gamma, bases, shifted, shifted_gamma, bounds, max_degree)

    def powers(self):
        return self.powers_of_beta_g, self.powers_of_beta_times_gamma_g

    def shifted_powers(self, degree_bound: int):
        """data_structures.rs:310-331 → (powers, gamma powers) of the shift that enforces `degree_bound`"""
        if self.shifted_powers_of_beta_g is None or degree_bound not in (self.enforced_degree_bounds or []):
            raise ValueError(f"degree bound {degree_bound} is not enforced by this committer key")
        return self.shifted_powers_of_beta_g[self.enforced_degree_bounds[-1] - degree_bound:], self.shifted_powers_of_beta_times_gamma_g[degree_bound]

    def lagrange_basis(self, size: int):
        if size not in self.lagrange_bases_at_beta_g:
            raise ValueError(f"UnsupportedLagrangeBasisSize({size})")
        return self.lagrange_bases_at_beta_g[size], self.powers_of_beta_times_gamma_g


def _check_degrees_and_bounds(ck: CommitterKey, p: LabeledPolynomial) -> None:
    """kzg10/mod.rs check_degrees_and_bounds: a bounded polynomial needs degree ≤ bound ≤ max_degree and an enforced bound"""
    if p.degree_bound is not None:
        if ck.enforced_degree_bounds is None or p.degree_bound not in ck.enforced_degree_bounds:
            raise ValueError(f"UnsupportedDegreeBound({p.degree_bound})")
        if p.polynomial.shape[0] - 1 > p.degree_bound or p.degree_bound > ck.max_degree:
            raise ValueError(f"IncorrectDegreeBound for {p.label}")


class SonicKZG10:
    @staticmethod
    def commit(ck: CommitterKey, polynomials: list, blindings: list | None = None):
        """mod.rs:177-257 → (commitments uint64[count, 18], [Randomness]).  All MSMs of the round — plain powers, shifted powers,
        Lagrange bases, blinding terms — go through ONE device pass (device.sonic_commit_batch).  `blindings[i]`: the blinding
        polynomial of polynomial i (hiding_bound + 2 Montgomery coefficients) when it is hiding, else None."""
        count = len(polynomials)
        blindings = list(blindings) if blindings is not None else [None] * count
        bases, gammas, polys, rands = [], [], [], []
        for p, b in zip(polynomials, blindings):
            _check_degrees_and_bounds(ck, p)
            if p.lagrange:
                n = p.polynomial.shape[0]
                size = 1 << max(n - 1, 0).bit_length()
                basis, gamma = ck.lagrange_basis(size)
                if n == 0 or size != basis.shape[0]:
                    raise ValueError("evaluations do not match the Lagrange basis size")
            elif p.degree_bound is not None:
                basis, gamma = ck.shifted_powers(p.degree_bound)
            else:
                basis, gamma = ck.powers()
            if p.hiding_bound is not None:
                if b is None or b.shape[0] != p.hiding_bound + 2:
                    raise ValueError(f"{p.label}: a hiding commitment needs a blinding polynomial of hiding_bound + 2 coefficients")
                rands.append(Randomness(b))
            else:
                b = None
                rands.append(Randomness())
            bases.append(basis); gammas.append(gamma); polys.append(p.polynomial)
        any_hiding = any(r.is_hiding() for r in rands)
        out = device.sonic_commit_batch(bases, polys, gammas if any_hiding else None,
                                        [r.blinding_polynomial for r in rands] if any_hiding else None)
        return out, rands

    @staticmethod
    def combine_for_open(ck: CommitterKey, labeled_polynomials: list, rands: list, challenges):
        """mod.rs:259-284 + combine_polynomials :548-565: Σ challenge_i·p_i with one squeezed challenge per polynomial"""
        poly, rand = None, Randomness()
        for p, r in zip(labeled_polynomials, rands):
            _check_degrees_and_bounds(ck, p)
            ch = int(next(challenges)) % _R_MOD
            poly = poly_axpy(poly, ch, p.polynomial)
            rand = rand.axpy(ch, r)
        return (poly if poly is not None else _zeros(0, ck.powers_of_beta_g.device)), rand

    @staticmethod
    def batch_open(ck: CommitterKey, labeled_polynomials: list, query_set: list, rands: list, challenges):
        """mod.rs:286-342 → [(w uint64[18], random_v uint64[4] or None)] in the order of the point names (a BTreeMap in the reference).
        query_set: iterable of (label, (point_name, point as int)).  `challenges` yields, per point, one value per polynomial opened
        there (labels in sorted order) and then the discarded `_randomizer`, exactly the reference's squeeze sequence."""
        poly_rand = {p.label: (p, r) for p, r in zip(labeled_polynomials, rands)}
        by_point: dict = {}
        for label, (point_name, point) in query_set:
            entry = by_point.setdefault(point_name, (point, set()))
            entry[1].add(label)
        powers, gamma = ck.powers()
        witnesses, hiding_witnesses, random_vs = [], [], []
        for point_name in sorted(by_point):
            point, labels = by_point[point_name]
            qp, qr = [], []
            for label in sorted(labels):
                if label not in poly_rand:
                    raise KeyError(f"MissingPolynomial {{ label: {label} }}")
                qp.append(poly_rand[label][0]); qr.append(poly_rand[label][1])
            polynomial, rand = SonicKZG10.combine_for_open(ck, qp, qr, challenges)
            next(challenges)                                                   # `_randomizer`
            z = _fr_int_to_mont(int(point) % _R_MOD)
            if polynomial.shape[0] > powers.shape[0]:
                raise ValueError("check_degree_is_too_large")
            # kzg10::open (mod.rs:303-321): the witness polynomials now, their commitments below — all points in ONE device pass
            w, rw = KZG10.compute_witness_polynomial(polynomial, z, rand.blinding_polynomial if rand.is_hiding() else None)
            witnesses.append(w); hiding_witnesses.append(rw)
            random_vs.append(device.poly_evaluate(rand.blinding_polynomial, z) if rand.is_hiding() else None)
        k = len(witnesses)
        any_hiding = any(h is not None for h in hiding_witnesses)
        ws = device.sonic_commit_batch([powers] * k, witnesses, [gamma] * k if any_hiding else None, hiding_witnesses if any_hiding else None)
        return [(ws[i], random_vs[i]) for i in range(k)]

    @staticmethod
    def open_combinations(ck: CommitterKey, linear_combinations: list, polynomials: list, rands: list, query_set: list, challenges):
        """mod.rs:413-475.  linear_combinations: [(lc_label, [(coeff as int, polynomial label or None for the constant term)])];
        the query set names LC labels.  Returns the BatchLCProof's list of (w, random_v)."""
        label_map = {p.label: (p, r) for p, r in zip(polynomials, rands)}
        lc_polys, lc_rands = [], []
        for lc_label, terms in linear_combinations:
            poly, rand = None, Randomness()
            degree_bound = hiding_bound = None
            num_polys = len(terms)
            for coeff, label in terms:
                if label is None:                                              # LCTerm::One: not committed, used by the verifier directly
                    continue
                if label not in label_map:
                    raise KeyError(f"MissingPolynomial {{ label: {label} }}")
                cur, cur_rand = label_map[label]
                if cur.degree_bound is not None:
                    if num_polys != 1:
                        raise ValueError(f"EquationHasDegreeBounds({lc_label})")
                    assert int(coeff) % _R_MOD == 1, "Coefficient must be one for degree-bounded equations"
                    assert degree_bound is None or degree_bound == cur.degree_bound
                    degree_bound = cur.degree_bound
                if cur.hiding_bound is not None:
                    hiding_bound = cur.hiding_bound if hiding_bound is None else max(hiding_bound, cur.hiding_bound)
                poly = poly_axpy(poly, int(coeff), cur.polynomial)
                rand = rand.axpy(int(coeff), cur_rand)
            dev = ck.powers_of_beta_g.device
            lc_polys.append(LabeledPolynomial(lc_label, poly if poly is not None else _zeros(0, dev), degree_bound, hiding_bound))
            lc_rands.append(rand)
        return SonicKZG10.batch_open(ck, lc_polys, query_set, lc_rands, challenges)


def synthetic_srs(max_degree: int, beta: int, gamma: int, dev="cuda"):
    """(powers_of_beta_g, powers_of_beta_times_gamma_g) = (β^i·G, γβ^i·G) for i ≤ max_degree (+1 for the gamma powers, which the
    reference keeps one degree further, mod.rs:104-105) — a universal setup with KNOWN trapdoor for tests and benches, where every
    commitment can be checked in the scalar field: commit(p, r) = (p(β) + γ·r(β))·G.  Built on the device from the generator by
    one fixed-base pass (device.generate_powers)."""
    return device.generate_powers(max_degree + 1, beta, 1, dev), device.generate_powers(max_degree + 2, beta, gamma, dev)

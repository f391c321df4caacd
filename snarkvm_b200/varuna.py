"""Varuna AHP prover rounds with every polynomial resident in HBM.

Device-side mirror of /root/reference/algorithms/src/snark/varuna for the part of `prove_batch` (varuna.rs:336-620) that is
bulk field arithmetic — the five `AHPForR1CS::prover_*_round` functions and the indexer's arithmetization:

    ahp/indexer/indexer.rs:121-200, ahp/matrices.rs:138-190, 239-254   Circuit            (domains, row / col / row_col_val on K, transposes)
    ahp/prover/round_functions/mod.rs:43-192, ahp/prover/state.rs:107-178   init_prover   (z_A, z_B, z_C by sparse mat-vec, x_poly)
    ahp/prover/round_functions/first.rs:129-160    prover_first_round   (w)
    ahp/prover/round_functions/third.rs:207-234    calculate_assignments (z)
    ahp/prover/round_functions/second.rs:77-146    prover_second_round  (h_0)
    ahp/prover/round_functions/third.rs:126-326    prover_third_round   (g_1, h_1, lineval sums)
    ahp/prover/round_functions/fourth.rs:151-245   prover_fourth_round  (g_a, g_b, g_c, sums)
    ahp/prover/round_functions/fifth.rs:41-67      prover_fifth_round   (h_2)
    ahp/selectors.rs:70-123                        apply_randomized_selector

for the NON-HIDING mode (VarunaNonHidingMode), one circuit, any batch of instances.  Everything O(n) runs in this library's
kernels (NTT passes, PolyMultiplier pipeline, divide_by_vanishing_poly, batch inversion, sparse mat-vec, elementwise Fr ops);
torch only owns the buffers and does index plumbing (gathers, concatenation).  Challenges are host scalars: the Fiat-Shamir
sponge (Poseidon) is sequential and stays on the CPU, so callers pass α, η, β, δ in.
Polynomials are CUDA tensors [m, 4] int64 (Montgomery Fr, low degree first, NOT trimmed: trailing zero coefficients may be present;
`trimmed()` gives the reference's canonical form on the host).
"""
from __future__ import annotations

import numpy as np
import torch

from . import device
from .algorithms import EvaluationDomain, _fr_int_to_mont, _fr_mont_to_int

R_MOD = 8444461749428370424248824938781546531375899335154063827935233455917409239041   # curves/src/bls12_377/fr.rs:138-145


def _mont(v: int) -> np.ndarray:
    return _fr_int_to_mont(v % R_MOD)


def _zeros(n: int, dev) -> torch.Tensor:
    return torch.zeros((n, 4), dtype=torch.int64, device=dev)


def _pad(x: torch.Tensor, n: int) -> torch.Tensor:
    if x.shape[0] == n:
        return x
    out = _zeros(n, x.device)
    out[: x.shape[0]] = x[:n]
    return out


def _scale(x: torch.Tensor, k: int) -> torch.Tensor:
    if x.shape[0] == 0 or k % R_MOD == 1:
        return x
    return device.fr_vec_op(x, _mont(k), device.FR_MUL)


def _add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a + b for polynomials of different lengths"""
    if a.shape[0] < b.shape[0]:
        a, b = b, a
    if b.shape[0] == 0:
        return a
    out = a.clone()
    device.fr_vec_op(out[: b.shape[0]], b, device.FR_ADD, out=out[: b.shape[0]])
    return out


def _sub(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    n = max(a.shape[0], b.shape[0])
    out = _pad(a, n).clone() if a.shape[0] == n else _pad(a, n)
    if b.shape[0]:
        device.fr_vec_op(out[: b.shape[0]], b, device.FR_SUB, out=out[: b.shape[0]])
    return out


def trimmed(poly: torch.Tensor) -> list:
    """host copy as canonical integers with trailing zeros removed (DensePolynomial::from_coefficients_vec, dense.rs:61-66)"""
    h = device.fr_from_mont(poly).cpu().numpy().view(np.uint64) if poly.shape[0] else np.zeros((0, 4), dtype=np.uint64)
    vals = [sum(int(v) << (64 * i) for i, v in enumerate(row)) for row in h]
    while vals and vals[-1] == 0:
        vals.pop()
    return vals


def polymul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """PolyMultiplier::multiply of two coefficient vectors (fft/polynomial/multiplier.rs:70-134)"""
    if a.shape[0] == 0 or b.shape[0] == 0:
        return _zeros(0, a.device)
    dom = EvaluationDomain.new(a.shape[0] + b.shape[0] - 1)
    return device.polymul([a.contiguous(), b.contiguous()], [], dom.log_size_of_group)


def divide_by_vanishing(p: torch.Tensor, domain: EvaluationDomain):
    return device.poly_divide_by_vanishing(p.contiguous(), domain.size)


def mul_by_vanishing(p: torch.Tensor, domain: EvaluationDomain) -> torch.Tensor:
    """DensePolynomial::mul_by_vanishing_poly (dense.rs:153-158): p·(x^n − 1)"""
    n, m = domain.size, p.shape[0]
    out = _zeros(n + m, p.device)
    out[n:] = p
    device.fr_vec_op(out[:m], p, device.FR_SUB, out=out[:m])       # out[:m] −= p (for m > n the shifted copy is already there)
    return out


def apply_randomized_selector(poly: torch.Tensor, combiner: int, target: EvaluationDomain, src: EvaluationDomain, remainder_witness: bool):
    """ahp/selectors.rs:70-123"""
    multiplier = combiner * src.size % R_MOD * pow(target.size, -1, R_MOD) % R_MOD
    if not remainder_witness:
        h, _rem = divide_by_vanishing(poly, src)               # the reference asserts a zero remainder; parity tests check the result
        return _scale(h, multiplier), None
    poly = _scale(poly, multiplier)
    h, xg = divide_by_vanishing(poly, src)
    if target.size != src.size:
        xg = mul_by_vanishing(xg, target)
        xg, _rem = divide_by_vanishing(xg, src)
    return h, xg


def reindex_by_subdomain(variable_size: int, input_size: int, index: np.ndarray) -> np.ndarray:
    """EvaluationDomain::reindex_by_subdomain (fft/domain.rs:322-344), vectorised"""
    if variable_size <= input_size:
        raise ValueError("other.size() must be smaller than self.size()")
    period = variable_size // input_size
    index = np.asarray(index, dtype=np.int64)
    i = index - input_size
    x = period - 1
    return np.where(index < input_size, index * period, i + i // x + 1)


class Matrix:
    """A sparse R1CS matrix in CSR form on the device: row_ptr int32 [nrows + 1], cols int32 [nnz] (variable indices: public first,
    then private — into_matrix_helper, ahp/matrices.rs:39-63), vals [nnz, 4] int64 Montgomery."""

    def __init__(self, row_ptr: np.ndarray, cols: np.ndarray, vals_mont: np.ndarray, dev):
        self.row_ptr_h = np.ascontiguousarray(row_ptr, dtype=np.int64)
        self.cols_h = np.ascontiguousarray(cols, dtype=np.int64)
        self.nrows, self.nnz = len(row_ptr) - 1, len(cols)
        self.row_ptr = torch.from_numpy(self.row_ptr_h.astype(np.int32)).to(dev)
        self.cols = torch.from_numpy(self.cols_h.astype(np.int32)).to(dev)
        self.vals = torch.from_numpy(np.ascontiguousarray(vals_mont, dtype=np.uint64).reshape(-1, 4).view(np.int64)).to(dev)


class MatrixEvals:
    """ahp/matrices.rs:102-136: row, col, row_col_val evaluations on the non-zero domain K"""

    def __init__(self, row, col, row_col_val, domain):
        self.row, self.col, self.row_col_val, self.domain = row, col, row_col_val, domain


class Circuit:
    """AHPForR1CS::index_helper (ahp/indexer/indexer.rs:121-200) for the non-hiding mode: the caller has already padded the public
    variables to a power of two (pad_input_for_indexer_and_prover, ahp/matrices.rs:85-100)."""

    def __init__(self, a: Matrix, b: Matrix, c: Matrix, num_public: int, num_variables: int):
        dev = a.vals.device
        self.a, self.b, self.c = a, b, c
        self.num_public, self.num_variables, self.num_constraints = num_public, num_variables, a.nrows
        if num_public & (num_public - 1):
            raise ValueError("public variables must be padded to a power of two")
        self.constraint_domain = EvaluationDomain.new(self.num_constraints)
        self.variable_domain = EvaluationDomain.new(num_variables)
        self.input_domain = EvaluationDomain.new(num_public)
        self.non_zero_domains = [EvaluationDomain.new(m.nnz) for m in (a, b, c)]
        self.max_non_zero_domain = max(self.non_zero_domains, key=lambda d: d.size)
        r_el = device.domain_elements(self.constraint_domain.log_size_of_group, dev)
        c_el = device.domain_elements(self.variable_domain.log_size_of_group, dev)
        one = torch.from_numpy(_mont(1).view(np.int64)).to(dev)
        self.ariths, self.transposes = [], []
        for m, K in zip((a, b, c), self.non_zero_domains):
            entry_rows = np.repeat(np.arange(m.nrows, dtype=np.int64), np.diff(m.row_ptr_h))
            entry_cols = reindex_by_subdomain(self.variable_domain.size, self.input_domain.size, m.cols_h)
            # matrix_evals (ahp/matrices.rs:138-190)
            row = one.repeat(K.size, 1)
            col = one.repeat(K.size, 1)
            rcv = _zeros(K.size, dev)
            if m.nnz:
                row[: m.nnz] = r_el[torch.from_numpy(entry_rows).to(dev)]
                col[: m.nnz] = c_el[torch.from_numpy(entry_cols).to(dev)]
                rc = device.fr_vec_op(row[: m.nnz].contiguous(), col[: m.nnz].contiguous(), device.FR_MUL)
                rcv[: m.nnz] = device.fr_vec_op(m.vals, rc, device.FR_MUL)
            self.ariths.append(MatrixEvals(row, col, rcv, K))
            # transpose (ahp/matrices.rs:239-254) as CSR over the variable domain's indices; entries keep the row-major order
            order = np.argsort(entry_cols, kind="stable")
            counts = np.bincount(entry_cols, minlength=self.variable_domain.size)
            t_ptr = np.concatenate([[0], np.cumsum(counts)])
            t_vals = m.vals[torch.from_numpy(order).to(dev)] if m.nnz else m.vals
            tm = Matrix.__new__(Matrix)
            tm.nrows, tm.nnz = self.variable_domain.size, m.nnz
            tm.row_ptr = torch.from_numpy(t_ptr.astype(np.int32)).to(dev)
            tm.cols = torch.from_numpy(entry_rows[order].astype(np.int32)).to(dev)
            tm.vals = t_vals.contiguous()
            self.transposes.append(tm)


class Prover:
    """init_prover + State::initialize (round_functions/mod.rs:43-170, state.rs:107-178) and the five rounds.
    `assignments`: one CUDA tensor [num_variables, 4] per instance — padded public variables (first one is One) then private."""

    def __init__(self, circuit: Circuit, assignments: list):
        self.circuit = c = circuit
        self.batch = len(assignments)
        self.z = [z.contiguous() for z in assignments]
        for z in self.z:
            if z.shape[0] != c.num_variables:
                raise ValueError("instance does not match the index")                       # AHPError::InstanceDoesNotMatchIndex
        self.z_a = [device.sparse_matvec(c.a.row_ptr, c.a.cols, c.a.vals, z) for z in self.z]
        self.z_b = [device.sparse_matvec(c.b.row_ptr, c.b.cols, c.b.vals, z) for z in self.z]
        self.z_c = [device.sparse_matvec(c.c.row_ptr, c.c.cols, c.c.vals, z) for z in self.z]
        self.x_polys = [c.input_domain.ifft(z[: c.num_public]) for z in self.z]                 # state.rs:137-139

    # ---- round 1: calculate_w (first.rs:129-160) ----
    def first_round(self):
        c = self.circuit
        V, I = c.variable_domain, c.input_domain
        dev = self.z[0].device
        # the index plumbing of calculate_w depends only on the two domain sizes: built once per circuit and device, on the device
        # (building it with numpy and uploading it every proof was 15 of the round's 21 ms at 2^20 constraints)
        cache = c.__dict__.setdefault("_w_index", {})
        if dev not in cache:
            ratio = V.size // I.size
            k = torch.arange(V.size, dtype=torch.int64, device=dev)
            mask = (k % ratio) != 0
            cache[dev] = (k[mask].contiguous(), (k - torch.div(k, ratio, rounding_mode="floor") - 1)[mask].contiguous())
        keep, src = cache[dev]
        self.w_polys = []
        for z, x_poly in zip(self.z, self.x_polys):
            w_ext = _zeros(V.size - I.size, dev)
            prv = z[c.num_public:]
            w_ext[: prv.shape[0]] = prv
            x_evals = V.fft_in_place(_pad(x_poly, V.size).clone())
            evals = _zeros(V.size, dev)
            evals[keep] = device.fr_vec_op(w_ext[src].contiguous(), x_evals[keep].contiguous(), device.FR_SUB)
            w_poly, _rem = divide_by_vanishing(V.ifft_in_place(evals), I)
            self.w_polys.append(w_poly)
        return self.w_polys

    # ---- calculate_assignments (third.rs:207-234): z = w·v_I + x ----
    def assignments(self):
        I = self.circuit.input_domain
        self.z_polys = [_add(mul_by_vanishing(w, I), x) for w, x in zip(self.w_polys, self.x_polys)]
        return self.z_polys

    # ---- round 2: calculate_rowcheck_witness (second.rs:77-146) ----
    def second_round(self, circuit_combiner: int = 1, instance_combiners=None):
        c = self.circuit
        Rd = c.constraint_domain
        instance_combiners = instance_combiners or [1] * self.batch
        h_0 = _zeros(0, self.z[0].device)
        for comb, za, zb, zc in zip(instance_combiners, self.z_a, self.z_b, self.z_c):
            pa, pb, pc = (Rd.ifft_in_place(_pad(e, Rd.size).clone()) for e in (za, zb, zc))
            rowcheck = _sub(polymul(pa, pb), pc)
            h_i, _ = apply_randomized_selector(_scale(rowcheck, comb), circuit_combiner, Rd, Rd, False)
            h_0 = _add(h_0, h_i)
        self.h_0 = h_0
        return h_0

    # ---- evaluate_all_lagrange_coefficients (fft/domain.rs:258-292) on the device ----
    @staticmethod
    def lagrange_coefficients(domain: EvaluationDomain, tau: int, dev) -> torch.Tensor:
        return domain.evaluate_all_lagrange_coefficients(tau, dev)

    # ---- round 3: lineval sumcheck (third.rs:126-205, 266-326) ----
    def third_round(self, alpha: int, eta_b: int, eta_c: int, circuit_combiner: int = 1, instance_combiners=None):
        c = self.circuit
        Rd, V = c.constraint_domain, c.variable_domain
        dev = self.z[0].device
        instance_combiners = instance_combiners or [1] * self.batch
        l_at_alpha = self.lagrange_coefficients(Rd, alpha, dev)
        h_1, xg_1, sums = _zeros(0, dev), _zeros(0, dev), []
        m_polys = []
        for tm in c.transposes:
            m_evals = device.sparse_matvec(tm.row_ptr, tm.cols, tm.vals, l_at_alpha)            # M(α, c) for every c ∈ C
            m_polys.append(V.ifft_in_place(m_evals))
        for inst_comb, z_poly in zip(instance_combiners, self.z_polys):
            inst_sums = []
            for m_at_alpha, m_comb in zip(m_polys, (1, eta_b % R_MOD, eta_c % R_MOD)):
                z_m = polymul(m_at_alpha, z_poly)
                # Σ_{c ∈ C} z_m(c) = |C| · Σ_j coefficient_{j·|C|}
                picks = device.fr_from_mont(z_m[:: V.size].contiguous()).cpu().numpy().view(np.uint64)
                inst_sums.append(V.size * sum(sum(int(v) << (64 * i) for i, v in enumerate(row)) for row in picks) % R_MOD)
                combiner = circuit_combiner * inst_comb % R_MOD * m_comb % R_MOD
                h_i, xg_i = apply_randomized_selector(z_m, combiner, V, V, True)
                h_1, xg_1 = _add(h_1, h_i), _add(xg_1, xg_i)
            sums.append(inst_sums)
        if self.mask_poly is not None:                                  # third.rs:207-213 (hiding mode)
            h_mask, xg_mask = divide_by_vanishing(self.mask_poly, V)
            h_1, xg_1 = _add(h_1, h_mask), _add(xg_1, xg_mask)
        self.h_1, self.g_1, self.third_sums = h_1, xg_1[1:].contiguous(), sums
        return self.g_1, self.h_1

    # ---- AHPForR1CS::construct_linear_combinations (ahp/ahp.rs:172-389) + verifier_query_set, one circuit ----
    def polynomials(self) -> dict:
        """label → device polynomial, everything prove_batch hands to open_combinations (varuna.rs:509-517)"""
        out = {f"w_{j}": w for j, w in enumerate(self.w_polys)}
        if self.mask_poly is not None:
            out["mask_poly"] = self.mask_poly
        out.update({"h_0": self.h_0, "g_1": self.g_1, "h_1": self.h_1, "h_2": self.h_2})
        for m, g, a, b in zip("abc", self.gs, self.a_polys, self.b_polys):
            out[f"g_{m}"], out[f"a_poly_{m}"], out[f"b_poly_{m}"] = g, a, b
        return out

    @staticmethod
    def _eval(poly: torch.Tensor, point: int) -> int:
        return _fr_mont_to_int(device.poly_evaluate(poly.contiguous(), _mont(point))) if poly.shape[0] else 0

    def linear_combinations(self, alpha, eta_b, eta_c, beta, deltas, gamma, circuit_combiner: int = 1, instance_combiners=None):
        """→ (lcs, query_set) exactly as oracle/varuna.py Prover.linear_combinations: the coefficients are host integers (a few dozen
        field operations), the three evaluations they need — g_1(β), g_M(γ), x_j(β) — are device Horner passes."""
        c = self.circuit
        Rd, V, I, K = c.constraint_domain, c.variable_domain, c.input_domain, c.max_non_zero_domain
        vanish = lambda d, x: (pow(x, d.size, R_MOD) - 1) % R_MOD       # noqa: E731
        instance_combiners = instance_combiners or [1] * self.batch
        lcs = {}
        const = 0
        for comb, sums in zip(instance_combiners, self.third_sums):
            const = (const + comb * (sums[0] * sums[1] - sums[2])) % R_MOD
        lcs["rowcheck_zerocheck"] = [(circuit_combiner * const % R_MOD, None), ((-vanish(Rd, alpha)) % R_MOD, "h_0")]
        lcs["g_1"] = [(1, "g_1")]
        v_c_beta, v_x_beta = vanish(V, beta), vanish(I, beta)
        g_1_at_beta = self._eval(self.g_1, beta)
        doms = [a.domain for a in c.ariths]
        sums4 = [s * d.size % R_MOD for s, d in zip(self.fourth_sums, doms)]
        weight = (sums4[0] + sums4[1] * eta_b + sums4[2] * eta_c) % R_MOD
        lineval = [(1, "mask_poly")] if self.mask_poly is not None else []
        for j, comb in enumerate(instance_combiners):
            x_at_beta = self._eval(self.x_polys[j], beta)
            k = circuit_combiner * comb % R_MOD
            lineval.append((k * weight % R_MOD * x_at_beta % R_MOD, None))
            lineval.append((k * weight % R_MOD * v_x_beta % R_MOD, f"w_{j}"))
        batch_lineval_sum = circuit_combiner * sum(comb * (s[0] + eta_b * s[1] + eta_c * s[2]) for comb, s in zip(instance_combiners, self.third_sums)) % R_MOD \
            * pow(V.size, -1, R_MOD) % R_MOD
        lineval += [((-v_c_beta) % R_MOD, "h_1"), ((-beta * g_1_at_beta) % R_MOD, None), ((-batch_lineval_sum) % R_MOD, None)]
        lcs["lineval_sumcheck"] = lineval
        v_k_gamma = vanish(K, gamma)
        matrix = []
        for m, g, s, delta, dom in zip("abc", self.gs, self.fourth_sums, deltas, doms):
            lcs[f"g_{m}"] = [(1, f"g_{m}")]
            selector = v_k_gamma * dom.size % R_MOD * pow(vanish(dom, gamma) * K.size % R_MOD, -1, R_MOD) % R_MOD
            b_term = (gamma * self._eval(g, gamma) + s) % R_MOD
            matrix.append((delta * selector % R_MOD, f"a_poly_{m}"))
            matrix.append(((-delta * selector % R_MOD * b_term) % R_MOD, f"b_poly_{m}"))
        matrix.append(((-v_k_gamma) % R_MOD, "h_2"))
        lcs["matrix_sumcheck"] = matrix
        points = {"rowcheck_zerocheck": ("alpha", alpha), "g_1": ("beta", beta), "lineval_sumcheck": ("beta", beta),
                  "g_a": ("gamma", gamma), "g_b": ("gamma", gamma), "g_c": ("gamma", gamma), "matrix_sumcheck": ("gamma", gamma)}
        order = sorted(lcs)
        return [(k, lcs[k]) for k in order], [(k, points[k]) for k in order]

    mask_poly = None

    def set_mask_poly(self, h_1_mask_rand, g_1_mask_rand):
        """calculate_mask_poly (first.rs:102-127) for the hiding mode: rand(degree 3)·v_H on the variable domain plus rand(degree 5)
        with a zero constant term; the random coefficients (ints) are arguments.  The mask is a first-round oracle (committed without a
        degree or hiding bound, first.rs:54-56) and enters h_1 / g_1 in the third round."""
        assert len(h_1_mask_rand) == 4 and len(g_1_mask_rand) == 6
        n = self.circuit.variable_domain.size
        mask = [0] * (n + 4)
        for i, c in enumerate(h_1_mask_rand):
            mask[n + i] = (mask[n + i] + c) % R_MOD
            mask[i] = (mask[i] - c) % R_MOD
        for i, c in enumerate(g_1_mask_rand):
            if i:
                mask[i] = (mask[i] + c) % R_MOD
        dev = self.z[0].device
        idx = sorted(set(range(6)) | set(range(n, n + 4)))              # the (at most ten) non-zero coefficients
        vals = np.array([_mont(mask[i]) for i in idx], dtype=np.uint64).reshape(-1, 4)
        self.mask_poly = _zeros(n + 4, dev)
        self.mask_poly[torch.tensor(idx, dtype=torch.int64, device=dev)] = torch.from_numpy(vals.view(np.int64)).to(dev)
        return self.mask_poly

    # ---- round 4: matrix sumchecks (fourth.rs:151-245) ----
    def fourth_round(self, alpha: int, beta: int):
        c = self.circuit
        Rd, V = c.constraint_domain, c.variable_domain
        v_rc = (pow(alpha, Rd.size, R_MOD) - 1) * (pow(beta, V.size, R_MOD) - 1) % R_MOD
        rc_size = Rd.size * V.size % R_MOD
        consts = v_rc * pow(Rd.size, -1, R_MOD) % R_MOD * pow(V.size, -1, R_MOD) % R_MOD
        self.gs, self.lhs, self.fourth_sums, self.a_polys, self.b_polys = [], [], [], [], []
        for arith in c.ariths:
            K = arith.domain
            a_poly = K.ifft_in_place(_scale(arith.row_col_val, v_rc).clone())
            # (α − r)(β − c) = (r − α)(c − β): the common factor of b's evaluations and of the denominators
            ra = device.fr_vec_op(arith.row, _mont(alpha), device.FR_SUB)
            cb = device.fr_vec_op(arith.col, _mont(beta), device.FR_SUB)
            prod = device.fr_vec_op(ra, cb, device.FR_MUL)
            b_poly = K.ifft_in_place(_scale(prod, rc_size).clone())                          # |R||C|(αβ − βr − αc + rc)
            device.fr_batch_inversion_and_mul(prod, _mont(consts))                            # fields/src/lib.rs:78-129
            f = K.ifft_in_place(device.fr_vec_op(prod, arith.row_col_val, device.FR_MUL))
            h = _sub(a_poly, polymul(b_poly, f))
            lhs, _ = apply_randomized_selector(h, 1, c.max_non_zero_domain, K, False)
            self.gs.append(f[1:].contiguous()); self.lhs.append(lhs)
            self.fourth_sums.append(_fr_mont_to_int(f[0].cpu().numpy().view(np.uint64)))
            self.a_polys.append(a_poly); self.b_polys.append(b_poly)
        return self.gs

    # ---- round 5 (fifth.rs:41-67) ----
    def fifth_round(self, deltas):
        h_2 = _zeros(0, self.z[0].device)
        for d, lhs in zip(deltas, self.lhs):
            h_2 = _add(h_2, _scale(lhs, d % R_MOD))
        self.h_2 = h_2
        return h_2

    def oracles(self) -> dict:
        """every polynomial the prover commits to, by round (varuna.rs:387-506)"""
        first = list(self.w_polys) + ([self.mask_poly] if self.mask_poly is not None else [])
        return {1: first, 2: [self.h_0], 3: [self.g_1, self.h_1], 4: list(self.gs), 5: [self.h_2]}


def test_circuit_csr(a: int, b: int, mul_depth: int, num_constraints: int, num_variables: int, dev):
    """TestCircuit (data_structures/test_circuit.rs:42-139) laid out directly as CSR matrices and one assignment, with the public
    variables padded to a power of two: every A, B, C row has a single unit entry.  Returns (Circuit, assignment tensor)."""
    num_public = 1 + mul_depth
    padded = 1 << (num_public - 1).bit_length() if num_public > 1 else 1
    num_private = num_variables - num_public
    nv = padded + num_private
    va, vb = padded + 0, padded + 1                                   # private variables a, b
    mul = [1 + i for i in range(mul_depth)]                            # public mul_var i
    mul_constraints = mul_depth - 1
    plain = num_constraints - mul_constraints
    a_cols = np.array([va] * plain + [mul[i] for i in range(mul_constraints)], dtype=np.int64)
    b_cols = np.full(num_constraints, vb, dtype=np.int64)
    c_cols = np.array([mul[0]] * plain + [mul[i + 1] for i in range(mul_constraints)], dtype=np.int64)
    row_ptr = np.arange(num_constraints + 1, dtype=np.int64)
    ones = np.tile(_mont(1), (num_constraints, 1))
    mats = [Matrix(row_ptr, cols, ones, dev) for cols in (a_cols, b_cols, c_cols)]
    circuit = Circuit(mats[0], mats[1], mats[2], padded, nv)
    z = np.zeros((nv, 4), dtype=np.uint64)
    z[0] = _mont(1)
    v = a % R_MOD
    for i in range(mul_depth):
        v = v * b % R_MOD
        z[1 + i] = _mont(v)
    z[padded:] = _mont(a)
    z[vb] = _mont(b)
    return circuit, torch.from_numpy(z.view(np.int64)).to(dev)

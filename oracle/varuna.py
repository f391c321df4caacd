"""TEST INFRASTRUCTURE ONLY — CPU restatement (Python big integers) of the Varuna AHP prover rounds.

Restates, function by function, /root/reference/algorithms/src/snark/varuna:
    data_structures/test_circuit.rs:42-90      TestCircuit::generate_constraints            → test_circuit()
    ahp/matrices.rs:85-100, 138-190, 239-254   pad_input…, matrix_evals, transpose          → pad / matrix_evals / transpose
    ahp/indexer/indexer.rs:121-200             index_helper (domains, arithmetization)      → Circuit
    ahp/prover/round_functions/mod.rs:43-192   init_prover + inner_product                  → Prover.__init__
    ahp/prover/state.rs:107-178                State::initialize (x_poly)                   → Prover.__init__
    ahp/prover/round_functions/first.rs:129-160   calculate_w                               → first_round
    ahp/prover/round_functions/third.rs:207-234   calculate_assignments                     → assignments
    ahp/prover/round_functions/second.rs:77-146   calculate_rowcheck_witness                → second_round
    ahp/selectors.rs:70-123                    apply_randomized_selector                    → apply_randomized_selector
    ahp/prover/round_functions/third.rs:126-205, 266-326   lineval sumcheck                 → third_round
    ahp/prover/round_functions/fourth.rs:151-245  calculate_matrix_sumcheck_witness         → fourth_round
    ahp/prover/round_functions/fifth.rs:41-67    h_2                                        → fifth_round
for the NON-HIDING mode (VarunaNonHidingMode: no randomizing variables, no mask polynomial, zk_bound = None), one
circuit, any number of instances.  Values are canonical integers mod r (not Montgomery); polynomials are coefficient
lists, low degree first, trimmed of trailing zeros like DensePolynomial::from_coefficients_vec (dense.rs:61-66).

PINNED against the reference's own golden vectors — resources/circuit_0/{polynomials,domain}/*.txt produced by
test_varuna_with_prover_test_vectors (snark/varuna/tests.rs:623-803) — in tests/test_varuna_golden.py.  Those nine
vectors are what pins the iFFT / FFT / PolyMultiplier / divide_by_vanishing_poly / batch_inversion_and_mul /
evaluate_all_lagrange_coefficients restatements used by the f1–f3 parity tests.
"""
from __future__ import annotations

from . import bls12_377 as py

R = py.R_MOD


# --------------------------------------------------------------------------------------------
# polynomials and domains
# --------------------------------------------------------------------------------------------
def trim(c):
    c = list(c)
    while c and c[-1] == 0:
        c.pop()
    return c


def poly_add(a, b):
    n = max(len(a), len(b))
    return trim([((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % R for i in range(n)])


def poly_sub(a, b):
    n = max(len(a), len(b))
    return trim([((a[i] if i < len(a) else 0) - (b[i] if i < len(b) else 0)) % R for i in range(n)])


def poly_scale(a, k):
    return trim([x * k % R for x in a])


def poly_eval(a, z):
    acc = 0
    for c in reversed(a):
        acc = (acc * z + c) % R
    return acc


class Domain:
    """EvaluationDomain::new (fft/domain.rs:118-147)"""

    def __init__(self, num_coeffs: int):
        size = 1
        while size < max(num_coeffs, 1):
            size <<= 1
        self.size = size
        self.lg = size.bit_length() - 1
        self.gen = py.fr_root_of_unity(size) if size > 1 else 1
        self.size_inv = pow(size, -1, R)

    def elements(self):
        out, x = [], 1
        for _ in range(self.size):
            out.append(x)
            x = x * self.gen % R
        return out

    def fft(self, coeffs):
        """evaluate over the domain; longer inputs are folded modulo x^n − 1 first (Polynomial::evaluate_over_domain,
        fft/polynomial/mod.rs:276-301)"""
        c = [0] * self.size
        for i, v in enumerate(coeffs):
            c[i % self.size] = (c[i % self.size] + v) % R
        return py.fft(c)

    def ifft(self, evals):
        e = list(evals) + [0] * (self.size - len(evals))
        assert len(e) == self.size
        return trim(py.ifft(e))

    def evaluate_vanishing_polynomial(self, tau):
        return (pow(tau, self.size, R) - 1) % R

    def evaluate_all_lagrange_coefficients(self, tau):
        """fft/domain.rs:258-292"""
        t_size = pow(tau, self.size, R)
        if t_size == 1:
            u, w = [0] * self.size, 1
            for i in range(self.size):
                if w == tau:
                    u[i] = 1
                    break
                w = w * self.gen % R
            return u
        l = (t_size - 1) * self.size_inv % R
        r = 1
        out = []
        for _ in range(self.size):
            out.append(l * pow((tau - r) % R, -1, R) % R)
            l = l * self.gen % R
            r = r * self.gen % R
        return out

    def reindex_by_subdomain(self, other: "Domain", index: int) -> int:
        """fft/domain.rs:322-344"""
        assert self.size > other.size
        period = self.size // other.size
        if index < other.size:
            return index * period
        i = index - other.size
        x = period - 1
        return i + (i // x) + 1


def poly_mul(a, b):
    """PolyMultiplier::multiply of two polynomials (fft/polynomial/multiplier.rs:70-134): FFT on the domain of
    next_pow2(deg sum + 1)"""
    if not a or not b:
        return []
    d = Domain(len(a) + len(b) - 1)
    ea, eb = d.fft(a), d.fft(b)
    return d.ifft([x * y % R for x, y in zip(ea, eb)])


def divide_by_vanishing_poly(p, domain: Domain):
    """DensePolynomial::divide_by_vanishing_poly (dense.rs:162-169) → (quotient, remainder), by the definition
    p = q·(x^n − 1) + r with deg r < n"""
    n = domain.size
    p = list(p)
    if len(p) <= n:
        return [], trim(p)
    q = [0] * (len(p) - n)
    for i in range(len(p) - 1, n - 1, -1):           # peel x^i = x^{i-n}·(x^n − 1) + x^{i-n}
        c = p[i]
        q[i - n] = c
        p[i - n] = (p[i - n] + c) % R
        p[i] = 0
    return trim(q), trim(p[:n])


def mul_by_vanishing_poly(p, domain: Domain):
    """dense.rs:153-158"""
    out = [0] * domain.size + list(p)
    for i, c in enumerate(p):
        out[i] = (out[i] - c) % R
    return trim(out)


def apply_randomized_selector(poly, combiner, target: Domain, src: Domain, remainder_witness: bool):
    """ahp/selectors.rs:70-123"""
    multiplier = combiner * src.size % R * target.size_inv % R
    if not remainder_witness:
        h, rem = divide_by_vanishing_poly(poly, src)
        assert rem == [], "non-zero remainder"
        return poly_scale(h, multiplier), None
    poly = poly_scale(poly, multiplier)
    h, xg = divide_by_vanishing_poly(poly, src)
    xg = mul_by_vanishing_poly(xg, target)
    xg, rem = divide_by_vanishing_poly(xg, src)
    assert rem == [], "non-zero remainder"
    return h, xg


# --------------------------------------------------------------------------------------------
# constraint system, test circuit, indexer
# --------------------------------------------------------------------------------------------
class ConstraintSystem:
    """public variable 0 is One; alloc_input → public, alloc → private; enforce(a, b, c) appends one row to A, B, C.
    Linear combinations are lists of (coefficient, ('pub' | 'priv', index))."""

    def __init__(self):
        self.public, self.private = [1], []
        self.a, self.b, self.c = [], [], []

    def alloc(self, v):
        self.private.append(v % R)
        return ("priv", len(self.private) - 1)

    def alloc_input(self, v):
        self.public.append(v % R)
        return ("pub", len(self.public) - 1)

    def enforce(self, a, b, c):
        self.a.append(list(a)); self.b.append(list(b)); self.c.append(list(c))


def test_circuit(a: int, b: int, mul_depth: int, num_constraints: int, num_variables: int) -> ConstraintSystem:
    """TestCircuit::generate_constraints (data_structures/test_circuit.rs:42-90)"""
    cs = ConstraintSystem()
    va, vb = cs.alloc(a), cs.alloc(b)
    mul_vars = []
    for i in range(mul_depth):
        v = a
        for _ in range(1 + i):
            v = v * b % R
        mul_vars.append(cs.alloc_input(v))
    for _ in range(num_variables - 3 - mul_depth):
        cs.alloc(a)
    mul_constraints = mul_depth - 1
    for _ in range(num_constraints - mul_constraints):
        cs.enforce([(1, va)], [(1, vb)], [(1, mul_vars[0])])
    for i in range(mul_constraints):
        cs.enforce([(1, mul_vars[i])], [(1, vb)], [(1, mul_vars[i + 1])])
    assert len(cs.a) == num_constraints and len(cs.public) + len(cs.private) == num_variables
    return cs


def pad_input_for_indexer_and_prover(cs: ConstraintSystem):
    """ahp/matrices.rs:85-100"""
    padded = Domain(len(cs.public)).size
    while len(cs.public) < padded:
        cs.alloc_input(0)


def into_matrix(rows, num_public):
    """into_matrix_helper (ahp/matrices.rs:39-63): columns = public index, or num_public + private index; sorted, merged"""
    out = []
    for row in rows:
        m = {}
        for val, (kind, i) in row:
            col = i if kind == "pub" else num_public + i
            m[col] = (m.get(col, 0) + val) % R
        out.append([(m[c], c) for c in sorted(m)])
    return out


class MatrixEvals:
    def __init__(self, row, col, row_col_val, domain):
        self.row, self.col, self.row_col_val, self.domain = row, col, row_col_val, domain


def matrix_evals(matrix, non_zero_domain: Domain, variable_domain: Domain, input_domain: Domain, r_elems, c_elems):
    """ahp/matrices.rs:138-190"""
    rows, cols, rcv = [], [], []
    for row_index, row in enumerate(matrix):
        for val, var_index in row:
            r_i = r_elems[row_index]
            c_i = c_elems[variable_domain.reindex_by_subdomain(input_domain, var_index)]
            rows.append(r_i); cols.append(c_i); rcv.append(val * r_i % R * c_i % R)
    pad = non_zero_domain.size - len(rows)
    rows += [1] * pad; cols += [1] * pad; rcv += [0] * pad
    return MatrixEvals(rows, cols, rcv, non_zero_domain)


def transpose(matrix, variable_domain: Domain, input_domain: Domain):
    """ahp/matrices.rs:239-254"""
    t = [[] for _ in range(variable_domain.size)]
    for row_index, row in enumerate(matrix):
        for val, var_index in row:
            t[variable_domain.reindex_by_subdomain(input_domain, var_index)].append((val, row_index))
    return t


class Circuit:
    """AHPForR1CS::index_helper (ahp/indexer/indexer.rs:121-200) for the non-hiding mode"""

    def __init__(self, cs: ConstraintSystem):
        pad_input_for_indexer_and_prover(cs)
        self.num_public = len(cs.public)
        self.num_variables = len(cs.public) + len(cs.private)
        self.num_constraints = len(cs.a)
        self.a, self.b, self.c = (into_matrix(m, self.num_public) for m in (cs.a, cs.b, cs.c))
        nnz = [sum(len(r) for r in m) for m in (self.a, self.b, self.c)]
        self.constraint_domain = Domain(self.num_constraints)
        self.variable_domain = Domain(self.num_variables)
        self.input_domain = Domain(self.num_public)
        self.non_zero_domains = [Domain(n) for n in nnz]
        self.max_non_zero_domain = max(self.non_zero_domains, key=lambda d: d.size)
        r_el, c_el = self.constraint_domain.elements(), self.variable_domain.elements()
        self.ariths = [matrix_evals(m, d, self.variable_domain, self.input_domain, r_el, c_el)
                       for m, d in zip((self.a, self.b, self.c), self.non_zero_domains)]


# --------------------------------------------------------------------------------------------
# prover
# --------------------------------------------------------------------------------------------
def inner_product(public, private, row, num_public):
    """ahp/prover/round_functions/mod.rs:172-192"""
    acc = 0
    for coeff, i in row:
        v = public[i] if i < num_public else private[i - num_public]
        acc = (acc + v * coeff) % R
    return acc


class Prover:
    """init_prover + State::initialize for one circuit and a batch of instances (each a ConstraintSystem with the same
    shape as the indexed one, already carrying its witness)."""

    def __init__(self, circuit: Circuit, instances):
        self.circuit = circuit
        self.public, self.private, self.z_a, self.z_b, self.z_c, self.x_polys = [], [], [], [], [], []
        for cs in instances:
            pad_input_for_indexer_and_prover(cs)
            assert len(cs.public) == circuit.num_public and len(cs.public) + len(cs.private) == circuit.num_variables
            pub, prv = list(cs.public), list(cs.private)
            self.public.append(pub); self.private.append(prv)
            for m, dst in ((circuit.a, self.z_a), (circuit.b, self.z_b), (circuit.c, self.z_c)):
                dst.append([inner_product(pub, prv, row, circuit.num_public) for row in m])
            self.x_polys.append(circuit.input_domain.ifft(pub))          # state.rs:137-139
        self.batch = len(instances)

    # ---- round 1 (first.rs:129-160) ----
    def first_round(self):
        c = self.circuit
        V, I = c.variable_domain, c.input_domain
        ratio = V.size // I.size
        self.w_polys = []
        for prv, x_poly in zip(self.private, self.x_polys):
            w_ext = list(prv) + [0] * (V.size - I.size - len(prv))
            x_evals = V.fft(x_poly)
            evals = [0 if k % ratio == 0 else (w_ext[k - (k // ratio) - 1] - x_evals[k]) % R for k in range(V.size)]
            w_poly, rem = divide_by_vanishing_poly(V.ifft(evals), I)
            assert rem == []
            self.w_polys.append(w_poly)
        return self.w_polys

    # ---- calculate_assignments (third.rs:207-234) ----
    def assignments(self):
        I = self.circuit.input_domain
        self.z_polys = [poly_add(mul_by_vanishing_poly(w, I), x) for w, x in zip(self.w_polys, self.x_polys)]
        return self.z_polys

    # ---- round 2 (second.rs:77-146) ----
    def second_round(self, circuit_combiner=1, instance_combiners=None):
        c = self.circuit
        Rd = c.constraint_domain
        instance_combiners = instance_combiners or [1] * self.batch
        h_0 = []
        for comb, za, zb, zc in zip(instance_combiners, self.z_a, self.z_b, self.z_c):
            pa, pb, pc = Rd.ifft(za), Rd.ifft(zb), Rd.ifft(zc)
            rowcheck = poly_sub(poly_mul(pa, pb), pc)
            h_i, _ = apply_randomized_selector(poly_scale(rowcheck, comb), circuit_combiner, Rd, Rd, False)
            h_0 = poly_add(h_0, h_i)
        self.h_0 = h_0
        return h_0

    # ---- round 3 (third.rs:126-205, 266-326) ----
    def third_round(self, alpha, eta_b, eta_c, circuit_combiner=1, instance_combiners=None):
        c = self.circuit
        Rd, V = c.constraint_domain, c.variable_domain
        instance_combiners = instance_combiners or [1] * self.batch
        l_at_alpha = Rd.evaluate_all_lagrange_coefficients(alpha)
        transposes = [transpose(m, V, c.input_domain) for m in (c.a, c.b, c.c)]
        h_1, xg_1, sums = [], [], []
        for inst_comb, z_poly in zip(instance_combiners, self.z_polys):
            inst_sums = []
            for mt, m_comb in zip(transposes, (1, eta_b, eta_c)):
                m_evals = [sum(val * l_at_alpha[row] for val, row in col) % R for col in mt]
                m_at_alpha = V.ifft(m_evals)
                z_m = poly_mul(m_at_alpha, z_poly)
                inst_sums.append(sum(V.fft(z_m)) % R)
                combiner = circuit_combiner * inst_comb % R * m_comb % R
                h_i, xg_i = apply_randomized_selector(z_m, combiner, V, V, True)
                h_1, xg_1 = poly_add(h_1, h_i), poly_add(xg_1, xg_i)
            sums.append(inst_sums)
        if self.mask_poly is not None:                                  # third.rs:207-213 (hiding mode)
            h_mask, xg_mask = divide_by_vanishing_poly(self.mask_poly, V)
            h_1, xg_1 = poly_add(h_1, h_mask), poly_add(xg_1, xg_mask)
        self.h_1, self.g_1, self.third_sums = h_1, trim(xg_1[1:]), sums
        return self.g_1, self.h_1

    # ---- AHPForR1CS::construct_linear_combinations (ahp/ahp.rs:172-389) + verifier_query_set, one circuit ----
    def polynomials(self):
        """label → coefficients of every polynomial prove_batch hands to open_combinations (varuna.rs:509-517)"""
        out = {f"w_{j}": w for j, w in enumerate(self.w_polys)}
        if self.mask_poly is not None:
            out["mask_poly"] = trim(self.mask_poly)
        out.update({"h_0": self.h_0, "g_1": self.g_1, "h_1": self.h_1, "h_2": self.h_2})
        for m, g, a, b in zip("abc", self.gs, self.a_polys, self.b_polys):
            out[f"g_{m}"], out[f"a_poly_{m}"], out[f"b_poly_{m}"] = g, a, b
        return out

    def linear_combinations(self, alpha, eta_b, eta_c, beta, deltas, gamma, circuit_combiner=1, instance_combiners=None):
        """→ (lcs, query_set): lcs = [(label, [(coefficient, polynomial label or None for LCTerm::One)])] in the reference's BTreeMap
        order, query_set = [(lc label, (point name, point))].  The three checks evaluate to ZERO at their points (the reference's
        debug_asserts, ahp.rs:258, 340, 384) — tests/test_varuna_golden.py verifies exactly that."""
        c = self.circuit
        Rd, V, I = c.constraint_domain, c.variable_domain, c.input_domain
        K = c.max_non_zero_domain
        polys = self.polynomials()
        instance_combiners = instance_combiners or [1] * self.batch
        lcs = {}
        # rowcheck_zerocheck at α (ahp.rs:230-256); the selector of the circuit's own (= max) constraint domain is 1
        const = 0
        for comb, sums in zip(instance_combiners, self.third_sums):
            const = (const + comb * (sums[0] * sums[1] - sums[2])) % R
        lcs["rowcheck_zerocheck"] = [(circuit_combiner * const % R, None), ((-Rd.evaluate_vanishing_polynomial(alpha)) % R, "h_0")]
        # g_1 and lineval_sumcheck at β (ahp.rs:262-343)
        lcs["g_1"] = [(1, "g_1")]
        v_c_beta, v_x_beta = V.evaluate_vanishing_polynomial(beta), I.evaluate_vanishing_polynomial(beta)
        g_1_at_beta = poly_eval(self.g_1, beta)
        sums4 = [s * d.size % R for s, d in zip(self.fourth_sums, c.non_zero_domains)]
        weight = (sums4[0] + sums4[1] * eta_b + sums4[2] * eta_c) % R
        lineval = [(1, "mask_poly")] if self.mask_poly is not None else []
        for j, comb in enumerate(instance_combiners):
            x_at_beta = poly_eval(self.x_polys[j], beta)
            k = circuit_combiner * comb % R
            lineval.append((k * weight % R * x_at_beta % R, None))
            lineval.append((k * weight % R * v_x_beta % R, f"w_{j}"))
        batch_lineval_sum = circuit_combiner * sum(comb * (s[0] + eta_b * s[1] + eta_c * s[2]) for comb, s in zip(instance_combiners, self.third_sums)) % R * V.size_inv % R
        lineval += [((-v_c_beta) % R, "h_1"), ((-beta * g_1_at_beta) % R, None), ((-batch_lineval_sum) % R, None)]
        lcs["lineval_sumcheck"] = lineval
        # g_a, g_b, g_c and matrix_sumcheck at γ (ahp.rs:345-386)
        v_k_gamma = K.evaluate_vanishing_polynomial(gamma)
        matrix = []
        for m, g, s, delta, dom in zip("abc", self.gs, self.fourth_sums, deltas, c.non_zero_domains):
            lcs[f"g_{m}"] = [(1, f"g_{m}")]
            selector = v_k_gamma * dom.size % R * pow(dom.evaluate_vanishing_polynomial(gamma) * K.size % R, -1, R) % R
            b_term = (gamma * poly_eval(g, gamma) + s) % R
            matrix.append((delta * selector % R, f"a_poly_{m}"))
            matrix.append(((-delta * selector % R * b_term) % R, f"b_poly_{m}"))
        matrix.append(((-v_k_gamma) % R, "h_2"))
        lcs["matrix_sumcheck"] = matrix
        points = {"rowcheck_zerocheck": ("alpha", alpha), "g_1": ("beta", beta), "lineval_sumcheck": ("beta", beta),
                  "g_a": ("gamma", gamma), "g_b": ("gamma", gamma), "g_c": ("gamma", gamma), "matrix_sumcheck": ("gamma", gamma)}
        order = sorted(lcs)
        return [(k, lcs[k]) for k in order], [(k, points[k]) for k in order]

    def evaluate_lc(self, terms, point):
        polys = self.polynomials()
        return sum(coeff * (1 if label is None else poly_eval(polys[label], point)) for coeff, label in terms) % R

    mask_poly = None

    def set_mask_poly(self, h_1_mask_rand, g_1_mask_rand):
        """calculate_mask_poly (first.rs:102-127): h_1_mask = rand(degree 3)·v_H on the (max) variable domain, g_1_mask = rand(degree 5)
        with its constant coefficient zeroed; the two random polynomials are arguments (DensePolynomial::rand draws them)."""
        assert len(h_1_mask_rand) == 4 and len(g_1_mask_rand) == 6
        n = self.circuit.variable_domain.size
        mask = [0] * (n + 4)
        for i, c in enumerate(h_1_mask_rand):
            mask[n + i] = (mask[n + i] + c) % R
            mask[i] = (mask[i] - c) % R
        for i, c in enumerate(g_1_mask_rand):
            if i:
                mask[i] = (mask[i] + c) % R
        self.mask_poly = mask
        return mask

    # ---- round 4 (fourth.rs:151-245) ----
    def fourth_round(self, alpha, beta):
        c = self.circuit
        Rd, V = c.constraint_domain, c.variable_domain
        v_rc = Rd.evaluate_vanishing_polynomial(alpha) * V.evaluate_vanishing_polynomial(beta) % R
        self.gs, self.lhs, self.fourth_sums, self.a_polys, self.b_polys = [], [], [], [], []
        for arith in c.ariths:
            K = arith.domain
            a_poly = K.ifft([v_rc * v % R for v in arith.row_col_val])
            ab = alpha * beta % R
            b_poly = K.ifft([Rd.size * V.size % R * ((ab - beta * r - alpha * cc + r * cc) % R) % R
                             for r, cc in zip(arith.row, arith.col)])
            consts = v_rc * Rd.size_inv % R * V.size_inv % R
            inv = [(alpha - r) * (beta - cc) % R for r, cc in zip(arith.row, arith.col)]
            inv = [0 if x == 0 else consts * pow(x, -1, R) % R for x in inv]         # batch_inversion_and_mul (fields/src/lib.rs:78-129)
            f = K.ifft([i * v % R for i, v in zip(inv, arith.row_col_val)])
            g = trim(f[1:])
            h = poly_sub(a_poly, poly_mul(b_poly, f))
            lhs, _ = apply_randomized_selector(h, 1, c.max_non_zero_domain, K, False)
            self.gs.append(g); self.lhs.append(lhs); self.fourth_sums.append(f[0] if f else 0)
            self.a_polys.append(a_poly); self.b_polys.append(b_poly)
        return self.gs

    # ---- round 5 (fifth.rs:41-67) ----
    def fifth_round(self, deltas):
        h_2 = []
        for d, lhs in zip(deltas, self.lhs):
            h_2 = poly_add(h_2, poly_scale(lhs, d))
        self.h_2 = h_2
        return h_2

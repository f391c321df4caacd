"""Host-side mirror of the reference operator surface for the hot path
(`snarkvm_algorithms::{msm::VariableBase, fft::EvaluationDomain, fft::PolyMultiplier}`), so the
parity tests read like the reference's own tests.  Every method lands in libsnarkvm_b200.so:
numpy (host) arrays go through the drop-in FFI symbols exactly as the Rust callers would
(variable_base/mod.rs:33-42, fft/domain.rs:375-438, multiplier.rs:79-95); torch CUDA tensors go
through the device-resident API.  There is no CPU implementation in this package.
"""
from __future__ import annotations

import numpy as np

from . import cuda
from .cuda import NTTDirection, NTTInputOutputOrder, NTTType

FR_TWO_ADICITY = 47   # curves/src/bls12_377/fr.rs:109


def _is_torch(x) -> bool:
    try:
        import torch
        return isinstance(x, torch.Tensor)
    except ImportError:   # pragma: no cover
        return False


class VariableBase:
    """algorithms/src/msm/variable_base/mod.rs:27-49"""

    @staticmethod
    def msm(bases, scalars):
        """Σ scalars[i]·bases[i] over the first len(scalars) bases → projective uint64[18] (normalised).

        The reference dispatches G1-377 to the GPU only when len > 1024 (mod.rs:35) and otherwise runs
        batched::msm on the CPU; this backend has no CPU path, so every size runs on the device."""
        # the reference dispatches on the point TYPE (TypeId, mod.rs:33,44): here the image size stands for it —
        # 104-byte Affine<G1> rows take the batched-affine path, 200-byte Affine<G2> rows the generic Pippenger (standard::msm)
        g2 = getattr(bases, "ndim", 0) == 2 and bases.shape[1] == 200
        if _is_torch(bases):
            from . import device
            return device.msm_g2(bases, scalars) if g2 else device.msm(bases, scalars)
        return cuda.msm_g2(bases, scalars) if g2 else cuda.msm(bases, scalars)


class EvaluationDomain:
    """algorithms/src/fft/domain.rs:82-221 (size, log_size_of_group; the field constants live on the device)."""

    def __init__(self, size: int):
        self.size = size
        self.log_size_of_group = size.bit_length() - 1

    @classmethod
    def new(cls, num_coeffs: int):
        """domain.rs:118-147: next power of two, None above the 2-adicity."""
        size = 1 if num_coeffs <= 1 else 1 << (num_coeffs - 1).bit_length()
        if size.bit_length() - 1 > FR_TWO_ADICITY:
            return None
        return cls(size)

    @staticmethod
    def compute_size_of_domain(num_coeffs: int):
        d = EvaluationDomain.new(num_coeffs)
        return None if d is None else d.size

    # -- helpers -------------------------------------------------------------------------------
    def _resize(self, coeffs):
        """`coeffs.resize(self.size(), T::zero())` (domain.rs:171,187,218)."""
        n = coeffs.shape[0]
        if n == self.size:
            return coeffs
        if n > self.size:
            raise ValueError("more coefficients than the domain size")
        if _is_torch(coeffs):
            import torch
            out = torch.zeros((self.size,) + tuple(coeffs.shape[1:]), dtype=coeffs.dtype, device=coeffs.device)
            out[:n] = coeffs
            return out
        out = np.zeros((self.size,) + coeffs.shape[1:], dtype=coeffs.dtype)
        out[:n] = coeffs
        return out

    def _run(self, x, direction: NTTDirection, ntt_type: NTTType):
        x = self._resize(x)
        if _is_torch(x):
            from . import device
            return device.ntt_(x, direction, ntt_type)
        x = np.ascontiguousarray(x)
        cuda.NTT(self.size, x, NTTInputOutputOrder.NN, direction, ntt_type)
        return x

    # -- the four public transforms; each returns the (possibly resized) transformed array -------
    def fft_in_place(self, coeffs):
        return self._run(coeffs, NTTDirection.Forward, NTTType.Standard)          # domain.rs:169-175

    def ifft_in_place(self, evals):
        return self._run(evals, NTTDirection.Inverse, NTTType.Standard)           # domain.rs:185-191

    def coset_fft_in_place(self, coeffs):
        return self._run(coeffs, NTTDirection.Forward, NTTType.Coset)             # domain.rs:201-206

    def coset_ifft_in_place(self, evals):
        return self._run(evals, NTTDirection.Inverse, NTTType.Coset)              # domain.rs:216-221

    def evaluate_all_lagrange_coefficients(self, tau: int, dev="cuda"):
        """fft/domain.rs:258-292 → CUDA tensor [size, 4] (Montgomery)"""
        return _lagrange_coefficients(self, tau % _R_MOD, dev)

    def fft(self, coeffs):
        return self.fft_in_place(coeffs.clone() if _is_torch(coeffs) else np.array(coeffs, copy=True))

    def ifft(self, evals):
        return self.ifft_in_place(evals.clone() if _is_torch(evals) else np.array(evals, copy=True))

    def coset_fft(self, coeffs):
        return self.coset_fft_in_place(coeffs.clone() if _is_torch(coeffs) else np.array(coeffs, copy=True))

    def coset_ifft(self, evals):
        return self.coset_ifft_in_place(evals.clone() if _is_torch(evals) else np.array(evals, copy=True))


def _lagrange_coefficients(domain: "EvaluationDomain", tau: int, dev):
    """EvaluationDomain::evaluate_all_lagrange_coefficients (fft/domain.rs:258-292) on the device → CUDA tensor [n, 4] Montgomery"""
    import torch
    from . import device
    n = domain.size
    t_size = pow(tau, n, _R_MOD)
    elems = device.domain_elements(domain.log_size_of_group, dev)
    one = torch.from_numpy(_fr_int_to_mont(1).view(np.int64)).to(dev)
    if t_size == 1:
        # tau is a domain element: the indicator vector of its position (domain.rs:264-275)
        tm = torch.from_numpy(_fr_int_to_mont(tau % _R_MOD).view(np.int64)).to(dev)
        out = torch.zeros((n, 4), dtype=torch.int64, device=dev)
        out[(elems == tm).all(dim=1)] = one
        return out
    # u_i = l·ω^i / (τ − ω^i), l = (τ^n − 1)/n: invert (ω^i − τ) with coefficient −l, multiply by ω^i (domain.rs:277-291)
    l = (t_size - 1) * pow(n, -1, _R_MOD) % _R_MOD
    u = device.fr_vec_op(elems, _fr_int_to_mont(tau % _R_MOD), device.FR_SUB)
    device.fr_batch_inversion_and_mul(u, _fr_int_to_mont((-l) % _R_MOD))
    return device.fr_vec_op(u, elems, device.FR_MUL)


class Evaluations:
    """fft/evaluations.rs:31-218 with the values resident in HBM: `evaluations` is a CUDA tensor [domain.size, 4] int64 (Montgomery
    Fr).  The reference's IFFTPrecomputation arguments have no counterpart: the twiddle table is cached per device inside the library."""

    def __init__(self, evaluations, domain: "EvaluationDomain"):
        self.evaluations, self._domain = evaluations, domain

    @classmethod
    def from_vec_and_domain(cls, evaluations, domain: "EvaluationDomain"):
        """evaluations.rs:40-43: resized (zero-padded or truncated) to the domain"""
        import torch
        n = domain.size
        if evaluations.shape[0] != n:
            e = torch.zeros((n, 4), dtype=evaluations.dtype, device=evaluations.device)
            k = min(n, evaluations.shape[0])
            e[:k] = evaluations[:k]
            evaluations = e
        return cls(evaluations.contiguous(), domain)

    def domain(self) -> "EvaluationDomain":
        return self._domain

    def interpolate_by_ref(self) -> "DensePolynomial":
        """evaluations.rs:46-48 (the input is kept)"""
        return DensePolynomial(self._domain.ifft_in_place(self.evaluations.clone()))

    def interpolate(self) -> "DensePolynomial":
        """evaluations.rs:59-63: consumes the evaluations (transformed in place)"""
        coeffs = self._domain.ifft_in_place(self.evaluations)
        self.evaluations = None
        return DensePolynomial(coeffs)

    def evaluate_with_coeffs(self, lagrange_coefficients_at_point):
        """evaluations.rs:91-93: Σ eval_i·L_i(point) → 32-byte Montgomery image (uint64[4])"""
        from . import device
        prod = device.fr_vec_op(self.evaluations, lagrange_coefficients_at_point, device.FR_MUL)
        return device.poly_evaluate(prod, _fr_int_to_mont(1))                      # Σ c_i·1^i

    def evaluate(self, point_mont):
        """evaluations.rs:86-89"""
        tau = _fr_mont_to_int(point_mont)
        return self.evaluate_with_coeffs(_lagrange_coefficients(self._domain, tau, self.evaluations.device))

    def _zip(self, other: "Evaluations", op):
        from . import device
        if self._domain.size != other._domain.size:
            raise ValueError("domains are unequal")                               # evaluations.rs:119,141,163,186
        return Evaluations(device.fr_vec_op(self.evaluations, other.evaluations, op), self._domain)

    def __mul__(self, other):
        from . import device
        return self._zip(other, device.FR_MUL)

    def __add__(self, other):
        from . import device
        return self._zip(other, device.FR_ADD)

    def __sub__(self, other):
        from . import device
        return self._zip(other, device.FR_SUB)

    def __truediv__(self, other):
        """evaluations.rs:181-189: elementwise a / b — one batched inversion of the divisor, then a product; a zero divisor
        panics in the reference (Field division) and raises here"""
        from . import device
        if self._domain.size != other._domain.size:
            raise ValueError("domains are unequal")
        if bool((other.evaluations == 0).all(dim=1).any()):
            raise ZeroDivisionError("division by a zero evaluation")
        inv = other.evaluations.clone()
        device.fr_batch_inversion_and_mul(inv, _fr_int_to_mont(1))
        return Evaluations(device.fr_vec_op(self.evaluations, inv, device.FR_MUL), self._domain)


class PolyMultiplier:
    """algorithms/src/fft/polynomial/multiplier.rs:28-134: collect polynomials / evaluations, multiply once."""

    def __init__(self):
        self.polynomials = []
        self.evaluations = []

    def add_polynomial(self, poly, _label: str = ""):
        self.polynomials.append(poly)

    def add_evaluation(self, evals, _label: str = ""):
        self.evaluations.append(evals)

    def multiply(self):
        """Returns the product's coefficients on the domain of size next_pow2(Σ len(poly)) — multiplier.rs:70-134;
        None when there is nothing to multiply or an evaluation's length differs from the domain."""
        if not self.polynomials and not self.evaluations:
            return None
        if self.polynomials:
            degree = sum(p.shape[0] for p in self.polynomials)
            domain = EvaluationDomain.new(degree)
        else:
            domain = EvaluationDomain.new(self.evaluations[0].shape[0])
        if domain is None or any(e.shape[0] != domain.size for e in self.evaluations):
            return None
        if _is_torch((self.polynomials + self.evaluations)[0]):
            from . import device
            return device.polymul(self.polynomials, self.evaluations, domain.log_size_of_group)
        return cuda.polymul(domain.size, [np.ascontiguousarray(p) for p in self.polynomials],
                            [np.ascontiguousarray(e) for e in self.evaluations])


_R_MOD = 8444461749428370424248824938781546531375899335154063827935233455917409239041   # curves/src/bls12_377/fr.rs:138-145


def _fr_mont_to_int(x) -> int:
    """host-side to_bigint of one Montgomery Fr (uint64[4]) — only for argument checks such as `point ∉ domain`"""
    limbs = np.ascontiguousarray(x, dtype=np.uint64).reshape(4)
    m = sum(int(v) << (64 * i) for i, v in enumerate(limbs))
    return m * pow(1 << 256, -1, _R_MOD) % _R_MOD


def _fr_int_to_mont(v: int) -> np.ndarray:
    m = (v << 256) % _R_MOD
    return np.array([(m >> (64 * i)) & (2**64 - 1) for i in range(4)], dtype=np.uint64)


class KZG10:
    """The MSM-bearing parts of KZG10 (algorithms/src/polycommit/kzg10/mod.rs:98-206) on device-resident operands
    (torch CUDA tensors; Montgomery coefficients, the reference's in-memory affine points)."""

    @staticmethod
    def commit(powers_of_beta_g, coefficients_mont, powers_of_beta_times_gamma_g=None, blinding_mont=None):
        """mod.rs:98-156: VariableBase::msm(powers, to_bigint(coeffs)) [+ msm(gamma powers, to_bigint(blinding)) when hiding;
        the caller samples the blinding polynomial]."""
        from . import device
        if blinding_mont is None:
            return device.kzg_commit(powers_of_beta_g, coefficients_mont)
        return device.kzg_commit_hiding(powers_of_beta_g, coefficients_mont, powers_of_beta_times_gamma_g, blinding_mont)

    @staticmethod
    def commit_lagrange(lagrange_basis_at_beta_g, evaluations_mont, powers_of_beta_times_gamma_g=None, blinding_mont=None):
        """mod.rs:159-206: the same MSM against the Lagrange basis; len(evaluations).next_power_of_two() must be the basis size."""
        from . import device
        n = evaluations_mont.shape[0]
        size = lagrange_basis_at_beta_g.numel() * lagrange_basis_at_beta_g.element_size() // device.AFFINE_STRIDE
        if n == 0 or (1 << (n - 1).bit_length()) != size:
            raise ValueError("evaluations do not match the Lagrange basis size")            # mod.rs:166-169
        return KZG10.commit(lagrange_basis_at_beta_g, evaluations_mont, powers_of_beta_times_gamma_g, blinding_mont)

    @staticmethod
    def compute_witness_polynomial(polynomial_mont, point_mont, blinding_mont=None):
        """mod.rs:220-241: (p / (x − point), blinding / (x − point) or None), quotients only"""
        from . import device
        w = device.poly_divide_by_linear(polynomial_mont, point_mont)
        rw = device.poly_divide_by_linear(blinding_mont, point_mont) if blinding_mont is not None else None
        return w, rw

    @staticmethod
    def open_with_witness_polynomial(powers_of_beta_g, point_mont, witness_mont, powers_of_beta_times_gamma_g=None,
                                     blinding_mont=None, hiding_witness_mont=None):
        """mod.rs:243-277 → (w as normalised projective uint64[18], random_v as Montgomery uint64[4] or None):
        w = msm(powers, witness) [+ msm(gamma powers, hiding witness)], random_v = blinding.evaluate(point)"""
        from . import device
        if hiding_witness_mont is None:
            return device.kzg_commit(powers_of_beta_g, witness_mont), None
        w = device.kzg_commit_hiding(powers_of_beta_g, witness_mont, powers_of_beta_times_gamma_g, hiding_witness_mont)
        return w, device.poly_evaluate(blinding_mont, point_mont)

    @staticmethod
    def open(powers_of_beta_g, polynomial_mont, point_mont, powers_of_beta_times_gamma_g=None, blinding_mont=None):
        """mod.rs:303-321: compute_witness_polynomial then open_with_witness_polynomial"""
        w, rw = KZG10.compute_witness_polynomial(polynomial_mont, point_mont, blinding_mont)
        return KZG10.open_with_witness_polynomial(powers_of_beta_g, point_mont, w, powers_of_beta_times_gamma_g, blinding_mont, rw)

    @staticmethod
    def open_lagrange(lagrange_basis_at_beta_g, domain: "EvaluationDomain", evaluations_mont, point_mont, evaluation_at_point_mont):
        """mod.rs:272-301: witness evaluations (eval_i − y)/(ω^i − point) committed against the Lagrange basis; the point must not
        lie in the domain and len(evaluations) must equal the domain (= basis) size."""
        from . import device
        n = evaluations_mont.shape[0]
        if (1 << max(n - 1, 0).bit_length()) != domain.size or n != domain.size:
            raise ValueError("`evaluations.len()` must equal `domain.size()`")                # mod.rs:286-290, 294
        z = _fr_mont_to_int(point_mont)
        if pow(z, domain.size, _R_MOD) == 1:
            raise ValueError("Point cannot be in the domain")                                  # mod.rs:283-285
        divisor = device.fr_vec_op(device.domain_elements(domain.log_size_of_group, evaluations_mont.device), point_mont, device.FR_SUB)
        device.fr_batch_inversion_and_mul(divisor, _fr_int_to_mont(1))                         # batch_inversion, mod.rs:293
        num = device.fr_vec_op(evaluations_mont, evaluation_at_point_mont, device.FR_SUB)
        device.fr_vec_op(divisor, num, device.FR_MUL, out=divisor)
        return KZG10.commit_lagrange(lagrange_basis_at_beta_g, divisor), None

    @staticmethod
    def batch_commit(powers_of_beta_g, polynomials_mont, powers_of_beta_times_gamma_g=None, blindings_mont=None):
        """all commitments of one round against the same powers in ONE device pass (sonic_pc/mod.rs:177-257) → [count, 18] uint64;
        `blindings_mont[i]` (or None) is polynomial i's blinding polynomial (hiding_bound = Some(_), kzg10/mod.rs:129-150)"""
        from . import device
        return device.kzg_commit_batch(powers_of_beta_g, list(polynomials_mont), gamma_powers=powers_of_beta_times_gamma_g,
                                       blindings_mont=None if blindings_mont is None else list(blindings_mont))


class UniversalParams:
    """polycommit/kzg10/data_structures.rs:36-100, the part that computes: lagrange_basis."""

    def __init__(self, powers_of_beta_g):
        self.powers_of_beta_g = powers_of_beta_g

    def lagrange_basis(self, domain: "EvaluationDomain"):
        """data_structures.rs:68-72: domain.ifft(powers_of_beta_g[0..domain.size]) normalised to affine"""
        from . import device
        nbytes = domain.size * device.AFFINE_STRIDE
        flat = self.powers_of_beta_g.reshape(-1).view(__import__("torch").uint8)
        if flat.numel() < nbytes:
            raise ValueError("not enough powers for this domain")
        return device.lagrange_basis(flat[:nbytes].contiguous())


class DensePolynomial:
    """The device-resident subset of fft/polynomial/dense.rs used between transforms: coefficients are a CUDA tensor [m, 4] int64
    (Montgomery Fr, low degree first)."""

    def __init__(self, coeffs):
        self.coeffs = coeffs

    def evaluate(self, point_mont):
        """dense.rs:98-114"""
        from . import device
        return device.poly_evaluate(self.coeffs, point_mont)

    def divide_by_vanishing_poly(self, domain: "EvaluationDomain"):
        """dense.rs:162-169 → (quotient, remainder) as DensePolynomial (not trimmed)"""
        from . import device
        q, r = device.poly_divide_by_vanishing(self.coeffs, domain.size)
        return DensePolynomial(q), DensePolynomial(r)

    def evaluate_over_domain(self, domain: "EvaluationDomain"):
        """dense.rs:172-181 → polynomial/mod.rs:266-290: zero-pad and FFT; when degree ≥ domain.size the reference FFTs every
        domain-sized chunk of coefficients and adds the results — the same values as one FFT of the coefficients folded
        modulo x^n − 1, which is the remainder of divide_by_vanishing_poly."""
        import torch
        from . import device
        coeffs = self.coeffs
        if coeffs.shape[0] > domain.size:
            _, coeffs = device.poly_divide_by_vanishing(coeffs, domain.size)
        x = torch.zeros((domain.size, 4), dtype=coeffs.dtype, device=coeffs.device)
        x[: coeffs.shape[0]] = coeffs
        return domain.fft_in_place(x)


def batch_inversion_and_mul(v, coeff_mont):
    """fields/src/lib.rs:78-129 on a CUDA tensor, in place"""
    from . import device
    return device.fr_batch_inversion_and_mul(v, coeff_mont)

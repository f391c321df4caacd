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
        if _is_torch(bases):
            from . import device
            return device.msm(bases, scalars)
        return cuda.msm(bases, scalars)


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

    def fft(self, coeffs):
        return self.fft_in_place(coeffs.clone() if _is_torch(coeffs) else np.array(coeffs, copy=True))

    def ifft(self, evals):
        return self.ifft_in_place(evals.clone() if _is_torch(evals) else np.array(evals, copy=True))

    def coset_fft(self, coeffs):
        return self.coset_fft_in_place(coeffs.clone() if _is_torch(coeffs) else np.array(coeffs, copy=True))

    def coset_ifft(self, evals):
        return self.coset_ifft_in_place(evals.clone() if _is_torch(evals) else np.array(evals, copy=True))


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


class KZG10:
    """The MSM-bearing core of KZG10::commit (algorithms/src/polycommit/kzg10/mod.rs:98-156)."""

    @staticmethod
    def commit(powers_of_beta_g, coefficients_mont):
        """commitment = VariableBase::msm(powers, to_bigint(coeffs)) with device-resident operands
        (torch CUDA tensors).  Hiding-polynomial randomness (mod.rs:123-156) is a second, 3-term MSM the
        caller adds; it is not on the hot path."""
        from . import device
        return device.kzg_commit(powers_of_beta_g, coefficients_mont)

"""TEST INFRASTRUCTURE — ctypes/numpy binding of oracle/liboracle.so (the C restatement
of the reference CPU path, oracle/oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package (snarkvm_b200/) never does.

Array conventions (identical to the reference's in-memory layouts):
  Fr vector      : np.uint64 [n, 4]   Montgomery limbs (fields/src/fp_256.rs:52)
  scalars        : np.uint64 [n, 4]   canonical integers < r (BigInteger256)
  affine bases   : np.uint8  [n, 104] x[48] y[48] inf[1] pad[7] (affine.rs:41-46)
  projective     : np.uint64 [18]     X, Y, Z Montgomery Fq (projective.rs:36-41)
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "liboracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_num_threads.restype = ctypes.c_int
        for name in ("oracle_ntt", "oracle_polymul", "oracle_msm", "oracle_fr_inverse", "oracle_fq_inverse",
                     "oracle_g1_is_on_curve"):
            getattr(_lib, name).restype = ctypes.c_int
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(t: int) -> None:
    lib().oracle_set_num_threads(ctypes.c_int(t))


def _binop(name, a, b, limbs):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(limbs)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(limbs)
    r = np.zeros(limbs, dtype=np.uint64)
    getattr(lib(), name)(_p(r), _p(a), _p(b))
    return r


def fr_mul(a, b): return _binop("oracle_fr_mul", a, b, 4)
def fr_add(a, b): return _binop("oracle_fr_add", a, b, 4)
def fr_sub(a, b): return _binop("oracle_fr_sub", a, b, 4)
def fq_mul(a, b): return _binop("oracle_fq_mul", a, b, 6)
def fq_add(a, b): return _binop("oracle_fq_add", a, b, 6)
def fq_sub(a, b): return _binop("oracle_fq_sub", a, b, 6)


def fr_inverse(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(4)
    r = np.zeros(4, dtype=np.uint64)
    ok = lib().oracle_fr_inverse(_p(r), _p(a))
    return r if ok else None


def fq_inverse(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(6)
    r = np.zeros(6, dtype=np.uint64)
    ok = lib().oracle_fq_inverse(_p(r), _p(a))
    return r if ok else None


def _conv(name, x, limbs):
    x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, limbs)
    out = np.empty_like(x)
    getattr(lib(), name)(_p(out), _p(x), ctypes.c_size_t(x.shape[0]))
    return out


def fr_to_mont(x): return _conv("oracle_fr_to_mont", x, 4)
def fr_from_mont(x): return _conv("oracle_fr_from_mont", x, 4)
def fq_to_mont(x): return _conv("oracle_fq_to_mont", x, 6)
def fq_from_mont(x): return _conv("oracle_fq_from_mont", x, 6)


FORWARD, INVERSE = 0, 1
STANDARD, COSET = 0, 1


def ntt(x: np.ndarray, direction: int = FORWARD, ntt_type: int = STANDARD) -> np.ndarray:
    """Out-of-place wrapper over the in-place oracle_ntt (NN order)."""
    x = np.array(x, dtype=np.uint64, order="C", copy=True).reshape(-1, 4)
    n = x.shape[0]
    lg = n.bit_length() - 1
    assert 1 << lg == n
    rc = lib().oracle_ntt(_p(x), ctypes.c_uint32(lg), 0, direction, ntt_type)
    assert rc == 0
    return x


def polymul(polys, evals, lg: int) -> np.ndarray:
    n = 1 << lg
    out = np.zeros((n, 4), dtype=np.uint64)
    polys = [np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 4) for p in polys]
    evals = [np.ascontiguousarray(e, dtype=np.uint64).reshape(-1, 4) for e in evals]
    pp = (ctypes.c_void_p * max(1, len(polys)))(*[p.ctypes.data for p in polys])
    pl = (ctypes.c_size_t * max(1, len(polys)))(*[p.shape[0] for p in polys])
    ep = (ctypes.c_void_p * max(1, len(evals)))(*[e.ctypes.data for e in evals])
    el = (ctypes.c_size_t * max(1, len(evals)))(*[e.shape[0] for e in evals])
    rc = lib().oracle_polymul(_p(out), ctypes.c_size_t(len(polys)), pp, pl, ctypes.c_size_t(len(evals)), ep, el,
                              ctypes.c_uint32(lg))
    assert rc == 0
    return out


BATCHED, STANDARD_MSM, NAIVE = 0, 1, 2


def msm(bases: np.ndarray, scalars: np.ndarray, algo: int = BATCHED, raw: bool = False) -> np.ndarray:
    """VariableBase::msm on the first len(scalars) bases → 18×u64 projective image
    (normalised: to_affine().to_projective(), unless raw)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    stride = bases.shape[1]
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    n = scalars.shape[0]
    assert n <= bases.shape[0]
    out = np.zeros(18, dtype=np.uint64)
    rc = lib().oracle_msm(_p(out), _p(bases), ctypes.c_size_t(n), _p(scalars), ctypes.c_size_t(stride),
                          ctypes.c_int(algo), ctypes.c_int(1 if raw else 0))
    assert rc == 0
    return out


def g1_normalise(p144) -> np.ndarray:
    p144 = np.ascontiguousarray(p144, dtype=np.uint64).reshape(18)
    out = np.zeros(18, dtype=np.uint64)
    lib().oracle_g1_normalise(_p(out), _p(p144))
    return out


def g1_mul(affine104, scalar) -> np.ndarray:
    a = np.ascontiguousarray(affine104, dtype=np.uint8).reshape(104)
    s = np.ascontiguousarray(scalar, dtype=np.uint64).reshape(4)
    out = np.zeros(18, dtype=np.uint64)
    lib().oracle_g1_mul(_p(out), _p(a), _p(s))
    return out


def g1_add(a144, b144) -> np.ndarray:
    a = np.ascontiguousarray(a144, dtype=np.uint64).reshape(18)
    b = np.ascontiguousarray(b144, dtype=np.uint64).reshape(18)
    out = np.zeros(18, dtype=np.uint64)
    lib().oracle_g1_add(_p(out), _p(a), _p(b))
    return out


def g1_is_on_curve(affine104) -> bool:
    a = np.ascontiguousarray(affine104, dtype=np.uint8).reshape(104)
    return bool(lib().oracle_g1_is_on_curve(_p(a)))


def batch_affine_add(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 104)
    b = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1, 104)
    out = np.zeros_like(a)
    lib().oracle_batch_affine_add(_p(out), _p(a), _p(b), ctypes.c_size_t(a.shape[0]))
    return out


def fr_dot_canonical(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Σ a_i·b_i mod r on canonical uint64 [n, 4] arrays → canonical uint64[4]."""
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    assert a.shape == b.shape
    out = np.zeros(4, dtype=np.uint64)
    lib().oracle_fr_dot_canonical(_p(out), _p(a), _p(b), ctypes.c_size_t(a.shape[0]))
    return out


# ---- next-row oracles (SURVEY §8 f1–f4) ----
def g1_ifft(bases: np.ndarray) -> np.ndarray:
    """UniversalParams::lagrange_basis (kzg10/data_structures.rs:68-72): group iFFT of n = 2^k affine points → affine [n, 104]."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1, 104)
    n = bases.shape[0]
    lg = n.bit_length() - 1
    assert 1 << lg == n
    out = np.zeros_like(bases)
    lib().oracle_g1_ifft.restype = ctypes.c_int
    assert lib().oracle_g1_ifft(_p(out), _p(bases), ctypes.c_uint32(lg)) == 0
    return out


def fr_batch_inversion_and_mul(v: np.ndarray, coeff: np.ndarray) -> np.ndarray:
    """fields/src/lib.rs:78-129 on Montgomery [n, 4] arrays (out of place); zeros stay zero."""
    v = np.array(v, dtype=np.uint64, order="C", copy=True).reshape(-1, 4)
    c = np.ascontiguousarray(coeff, dtype=np.uint64).reshape(4)
    lib().oracle_fr_batch_inversion_and_mul(_p(v), ctypes.c_size_t(v.shape[0]), _p(c))
    return v


def poly_divide_by_vanishing(p: np.ndarray, n: int):
    """DensePolynomial::divide_by_vanishing_poly (dense.rs:162-169): (q [max(m−n,0), 4], r [min(m,n), 4]), untrimmed."""
    p = np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 4)
    m = p.shape[0]
    q = np.zeros((max(m - n, 0), 4), dtype=np.uint64)
    r = np.zeros((min(m, n), 4), dtype=np.uint64)
    qd = np.zeros((max(q.shape[0], 1), 4), dtype=np.uint64)
    rd = np.zeros((max(r.shape[0], 1), 4), dtype=np.uint64)
    lib().oracle_poly_divide_by_vanishing.restype = ctypes.c_size_t
    lib().oracle_poly_divide_by_vanishing(_p(qd), _p(rd), _p(p), ctypes.c_size_t(m), ctypes.c_size_t(n))
    q[:] = qd[:q.shape[0]]
    r[:] = rd[:r.shape[0]]
    return q, r


def poly_evaluate(coeffs: np.ndarray, point: np.ndarray) -> np.ndarray:
    coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
    z = np.ascontiguousarray(point, dtype=np.uint64).reshape(4)
    out = np.zeros(4, dtype=np.uint64)
    lib().oracle_poly_evaluate(_p(out), _p(coeffs), ctypes.c_size_t(coeffs.shape[0]), _p(z))
    return out


def poly_divide_by_linear(p: np.ndarray, point: np.ndarray) -> np.ndarray:
    """compute_witness_polynomial (kzg10/mod.rs:220-241): quotient of p / (x − point) → [max(m − 1, 0), 4], untrimmed."""
    p = np.ascontiguousarray(p, dtype=np.uint64).reshape(-1, 4)
    z = np.ascontiguousarray(point, dtype=np.uint64).reshape(4)
    m = p.shape[0]
    q = np.zeros((max(m - 1, 1), 4), dtype=np.uint64)
    lib().oracle_poly_divide_by_linear.restype = ctypes.c_size_t
    lib().oracle_poly_divide_by_linear(_p(q), _p(p), ctypes.c_size_t(m), _p(z))
    return q[:max(m - 1, 0)]


def sparse_matvec(row_ptr, cols, vals, public_vars, private_vars) -> np.ndarray:
    """inner_product per row (varuna/ahp/prover/round_functions/mod.rs:169-189) for a CSR matrix → [nrows, 4]."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
    cols = np.ascontiguousarray(cols, dtype=np.uint32)
    vals = np.ascontiguousarray(vals, dtype=np.uint64).reshape(-1, 4)
    pub = np.ascontiguousarray(public_vars, dtype=np.uint64).reshape(-1, 4)
    prv = np.ascontiguousarray(private_vars, dtype=np.uint64).reshape(-1, 4)
    nrows = row_ptr.shape[0] - 1
    out = np.zeros((max(nrows, 1), 4), dtype=np.uint64)
    lib().oracle_sparse_matvec(_p(out), _p(row_ptr), _p(cols), _p(vals), ctypes.c_size_t(nrows), _p(pub), ctypes.c_size_t(pub.shape[0]), _p(prv))
    return out[:nrows]

"""Host mirror of the reference's Rust FFI shims — `snarkvm_algorithms_cuda::{NTT, polymul, msm}`
(/root/reference/algorithms/cuda/src/lib.rs:71-168) — over the same three C symbols.

Same names, argument meaning and error behaviour: caller bugs raise (the Rust shims `panic!`,
lib.rs:84-86,150-152), device failures raise CudaError (Rust returns `Err(cuda::Error)`).
Buffers are HOST numpy arrays in the reference's in-memory layouts.
"""
from __future__ import annotations

import ctypes
from enum import IntEnum

import numpy as np

from . import _lib


class NTTInputOutputOrder(IntEnum):   # lib.rs:22-28
    NN = 0
    NR = 1
    RN = 2
    RR = 3


class NTTDirection(IntEnum):          # lib.rs:30-34
    Forward = 0
    Inverse = 1


class NTTType(IntEnum):               # lib.rs:36-40
    Standard = 0
    Coset = 1


def _fr_array(x, name):
    if not (isinstance(x, np.ndarray) and x.dtype == np.uint64 and x.flags["C_CONTIGUOUS"] and x.ndim == 2 and x.shape[1] == 4):
        raise TypeError(f"{name} must be a C-contiguous uint64 array of shape [n, 4] (Montgomery Fr limbs)")
    return x


def NTT(domain_size: int, inout: np.ndarray, ntt_order: NTTInputOutputOrder, ntt_direction: NTTDirection,
        ntt_type: NTTType) -> None:
    """In-place NTT of `inout[..domain_size]` (lib.rs:77-97)."""
    if domain_size <= 0 or (domain_size & (domain_size - 1)) != 0:
        raise ValueError("domain_size is not power of 2")            # lib.rs:84-86
    _fr_array(inout, "inout")
    if inout.shape[0] < domain_size:
        raise ValueError("inout is shorter than domain_size")
    lg = domain_size.bit_length() - 1
    err = _lib.lib().snarkvm_ntt(inout.ctypes.data, lg, int(ntt_order), int(ntt_direction), int(ntt_type))
    _lib.check_rust_error(err)


def polymul(domain: int, polynomials: list, evaluations: list) -> np.ndarray:
    """Product of polynomials (coefficient form) and evaluations over `domain` (lib.rs:100-145)."""
    if domain <= 0 or (domain & (domain - 1)) != 0:
        raise ValueError("domain_size is not power of 2")            # lib.rs:108-110
    lg = domain.bit_length() - 1
    polys = [_fr_array(p, "polynomial") for p in polynomials]
    evals = [_fr_array(e, "evaluation") for e in evaluations]
    pp = (ctypes.c_void_p * max(1, len(polys)))(*[p.ctypes.data for p in polys])
    pl = (ctypes.c_size_t * max(1, len(polys)))(*[p.shape[0] for p in polys])
    ep = (ctypes.c_void_p * max(1, len(evals)))(*[e.ctypes.data for e in evals])
    el = (ctypes.c_size_t * max(1, len(evals)))(*[e.shape[0] for e in evals])
    out = np.zeros((domain, 4), dtype=np.uint64)                       # `out.resize(domain, zero)` lib.rs:126-127
    err = _lib.lib().snarkvm_polymul(out.ctypes.data, len(polys), ctypes.cast(pp, ctypes.c_void_p),
                                     ctypes.cast(pl, ctypes.c_void_p), len(evals), ctypes.cast(ep, ctypes.c_void_p),
                                     ctypes.cast(el, ctypes.c_void_p), lg)
    _lib.check_rust_error(err)
    return out


def msm(points: np.ndarray, scalars: np.ndarray) -> np.ndarray:
    """Σ scalars[i]·points[i] over the first len(scalars) points (lib.rs:148-168).

    points: uint8 [n, 104] (Affine<G1> images); scalars: uint64 [m, 4] canonical.  Returns the
    144-byte projective image as uint64[18] (normalised, Z = R)."""
    if not (isinstance(points, np.ndarray) and points.dtype == np.uint8 and points.ndim == 2 and points.flags["C_CONTIGUOUS"]):
        raise TypeError("points must be a C-contiguous uint8 array [n, ffi_affine_sz]")
    if not (isinstance(scalars, np.ndarray) and scalars.dtype == np.uint64 and scalars.ndim == 2 and scalars.shape[1] == 4
            and scalars.flags["C_CONTIGUOUS"]):
        raise TypeError("scalars must be a C-contiguous uint64 array [m, 4]")
    npoints = scalars.shape[0]
    if npoints > points.shape[0]:
        raise ValueError(f"length mismatch {points.shape[0]} points < {npoints} scalars")   # lib.rs:150-152
    out = np.zeros(18, dtype=np.uint64)
    err = _lib.lib().snarkvm_msm(out.ctypes.data, points.ctypes.data, npoints, scalars.ctypes.data, points.shape[1])
    _lib.check_rust_error(err)
    return out


def msm_g2(points: np.ndarray, scalars: np.ndarray) -> np.ndarray:
    """Σ scalars[i]·points[i] over G2: points uint8 [n, 200] (Affine<G2> images), scalars uint64 [m, 4] canonical → the 288-byte
    Projective<G2> image as uint64[36] (normalised).  Same length / error contract as `msm`."""
    if not (isinstance(points, np.ndarray) and points.dtype == np.uint8 and points.ndim == 2 and points.flags["C_CONTIGUOUS"]):
        raise TypeError("points must be a C-contiguous uint8 array [n, ffi_affine_sz]")
    if not (isinstance(scalars, np.ndarray) and scalars.dtype == np.uint64 and scalars.ndim == 2 and scalars.shape[1] == 4
            and scalars.flags["C_CONTIGUOUS"]):
        raise TypeError("scalars must be a C-contiguous uint64 array [m, 4]")
    npoints = scalars.shape[0]
    if npoints > points.shape[0]:
        raise ValueError(f"length mismatch {points.shape[0]} points < {npoints} scalars")
    out = np.zeros(36, dtype=np.uint64)
    err = _lib.lib().snarkvm_b200_msm_g2(out.ctypes.data, points.ctypes.data, npoints, scalars.ctypes.data, points.shape[1])
    _lib.check_rust_error(err)
    return out


def register_bases(points: np.ndarray) -> None:
    """Extension: keep `points` resident on the current device; later msm(points, …) calls with this same array skip
    the upload (the array must stay alive and unmodified until unregister_bases)."""
    if not (isinstance(points, np.ndarray) and points.dtype == np.uint8 and points.ndim == 2 and points.flags["C_CONTIGUOUS"]):
        raise TypeError("points must be a C-contiguous uint8 array [n, ffi_affine_sz]")
    _lib.check(_lib.lib().snarkvm_b200_register_bases(points.ctypes.data, points.shape[0], points.shape[1]))


def register_bases_precomputed(points: np.ndarray) -> None:
    """register_bases plus the fixed-base tables 2^{c·w}·P_i of the uploaded copy: `msm` on this array then runs over the tables."""
    if points.dtype != np.uint8 or points.ndim != 2 or not points.flags.c_contiguous:
        raise TypeError("points must be a C-contiguous uint8 array [n, stride]")
    _lib.check(_lib.lib().snarkvm_b200_register_bases_precomputed(points.ctypes.data, points.shape[0], points.shape[1]))


def unregister_bases(points: np.ndarray) -> None:
    _lib.check(_lib.lib().snarkvm_b200_unregister_bases(points.ctypes.data))

"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol the header declares,
and reports failure through the reference's error convention instead of aborting."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    with open(os.path.join(ROOT, "include", "snarkvm_b200.h")) as f:
        src = f.read()
    return sorted(set(re.findall(r"SNARKVM_API\s+[\w\s\*]+?\b(snarkvm_\w+)\s*\(", src)))


def test_header_declares_the_reference_ffi():
    syms = _header_symbols()
    for s in ("snarkvm_ntt", "snarkvm_polymul", "snarkvm_msm"):      # algorithms/cuda/src/lib.rs:42-69
        assert s in syms
    assert len(syms) >= 16


def test_library_exports_every_declared_symbol():
    import ctypes
    from snarkvm_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build with __graft_entry__.build()"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in _header_symbols():
        assert hasattr(L, s), s
    assert sorted(_lib.SYMBOLS) == _header_symbols()
    assert b"sm_100a" in _lib.lib().snarkvm_b200_version()


def test_only_snarkvm_symbols_are_exported():
    import subprocess
    from snarkvm_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    names = [l.split()[-1] for l in out.splitlines() if l.strip()]
    assert names and all(n.startswith("snarkvm_") for n in names), names


def test_error_convention_without_gpu():
    """No device ⇒ non-zero cudaError_t in .code, message malloc'd, outputs untouched, no abort
    (the Rust caller falls back to CPU on any non-zero code: variable_base/mod.rs:39-43)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from snarkvm_b200 import CudaError, cuda
    x = np.arange(32, dtype=np.uint64).reshape(8, 4)
    before = x.copy()
    with pytest.raises(CudaError) as ei:
        cuda.NTT(8, x, cuda.NTTInputOutputOrder.NN, cuda.NTTDirection.Forward, cuda.NTTType.Standard)
    assert ei.value.code != 0
    assert (x == before).all()
    with pytest.raises(CudaError):
        cuda.msm(np.zeros((4, 104), dtype=np.uint8), np.zeros((4, 4), dtype=np.uint64))


def test_shim_argument_checks():
    """The Rust shims panic on caller bugs (lib.rs:84-86, 150-152); the mirror raises ValueError."""
    from snarkvm_b200 import cuda
    x = np.zeros((6, 4), dtype=np.uint64)
    with pytest.raises(ValueError):
        cuda.NTT(6, x, cuda.NTTInputOutputOrder.NN, cuda.NTTDirection.Forward, cuda.NTTType.Standard)
    with pytest.raises(ValueError):
        cuda.msm(np.zeros((3, 104), dtype=np.uint8), np.zeros((4, 4), dtype=np.uint64))
    with pytest.raises(ValueError):
        cuda.polymul(12, [], [])


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under snarkvm_b200/ or include/ may reference it."""
    bad = []
    for d in ("snarkvm_b200", "include"):
        for root, _, files in os.walk(os.path.join(ROOT, d)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                    with open(os.path.join(root, fn)) as f:
                        for ln, line in enumerate(f, 1):
                            if re.search(r"^\s*(from|import)\s+oracle\b|#include\s+[\"<].*oracle", line):
                                bad.append(f"{fn}:{ln}")
    assert not bad, bad

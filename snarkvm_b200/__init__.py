"""snarkvm_b200 — B200 (sm_100a) backend for snarkVM's hot path: BLS12-377 G1 MSM and Fr NTT.

Layout: csrc/ (hand-written CUDA kernels + the C ABI of include/snarkvm_b200.h),
cuda.py (mirror of the reference's Rust FFI shims), algorithms.py (mirror of the reference operator
surface), device.py (HBM-resident torch entry points), sharded.py (multi-GPU MSM).
"""
from ._lib import CudaError, LIB_PATH, launch_count  # noqa: F401

__all__ = ["CudaError", "LIB_PATH", "launch_count"]

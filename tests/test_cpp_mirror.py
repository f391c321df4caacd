"""The C++ host-side mirror (include/snarkvm_b200.hpp) over the C ABI: compiles against the shared library (CPU test)
and, on a GPU, reproduces the oracle through VariableBase::msm / EvaluationDomain / PolyMultiplier."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "mirror_driver")
    libdir = os.path.join(ROOT, "snarkvm_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "mirror_driver.cpp"), "-o", exe,
                           "-L", libdir, "-lsnarkvm_b200", f"-Wl,-rpath,{libdir}"])
    return exe


def _write_inputs(tmp_path, cpu):
    from helpers import oracle_bases, random_canonical_fr, random_fr_mont
    n = 2000
    bases = oracle_bases(cpu, n, seed=5)
    scal = random_canonical_fr(n - 100, seed=6)          # fewer scalars than bases
    x = random_fr_mont(1 << 11, seed=7)
    p1, p2 = random_fr_mont(300, 8), random_fr_mont(500, 9)
    for name, arr in (("bases", bases), ("scalars", scal), ("fr", x), ("p1", p1), ("p2", p2)):
        arr.tofile(tmp_path / f"{name}.bin")
    return bases, scal, x, p1, p2


def test_cpp_mirror_compiles_and_reports_errors_without_gpu(tmp_path, oracle_cpu):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    _write_inputs(tmp_path, oracle_cpu)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 3 and "cuda error" in r.stderr      # Err(cuda::Error) surfaced, no abort, no CPU fallback


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle(tmp_path, oracle_cpu):
    exe = _build(tmp_path)
    bases, scal, x, p1, p2 = _write_inputs(tmp_path, oracle_cpu)
    # G2 inputs: 40 multiples of the generator (big-int oracle) and random scalars
    from oracle import g2
    from helpers import random_canonical_fr
    g2_pts = [g2.g2_mul(g2.G2_GEN, 1000 + 7 * i) for i in range(40)]
    g2_scal = random_canonical_fr(40, seed=66)
    np.frombuffer(b"".join(g2.g2_affine_bytes(q) for q in g2_pts), dtype=np.uint8).tofile(tmp_path / "g2_bases.bin")
    g2_scal.tofile(tmp_path / "g2_scalars.bin")
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got2 = np.fromfile(tmp_path / "msm_g2.out", dtype=np.uint64)
    ints = [sum(int(v) << (64 * i) for i, v in enumerate(row)) for row in g2_scal]
    assert (got2 == np.frombuffer(g2.g2_projective_bytes_normalised(g2.standard_msm(g2_pts, ints)), dtype=np.uint64)).all()
    got = np.fromfile(tmp_path / "msm.out", dtype=np.uint64)
    assert (got == oracle_cpu.msm(bases, scal, 0)).all()
    for name, d, t in (("fft", 0, 0), ("ifft", 1, 0), ("coset_fft", 0, 1), ("coset_ifft", 1, 1)):
        y = np.fromfile(tmp_path / f"{name}.out", dtype=np.uint64).reshape(-1, 4)
        assert (y == oracle_cpu.ntt(x, d, t)).all(), name
    prod = np.fromfile(tmp_path / "polymul.out", dtype=np.uint64).reshape(-1, 4)
    assert (prod == oracle_cpu.polymul([p1, p2], [], 10)).all()

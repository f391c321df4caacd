"""GPU parity for the G2 MSM (points over Fq2): the device Pippenger vs the big-int restatement of standard::msm
(algorithms/src/msm/variable_base/standard.rs:79-118) on the to_affine()-normalised Projective<G2> image, as the reference's own
test compares (msm/variable_base/mod.rs:90-119)."""
import random

import numpy as np
import pytest

from oracle import bls12_377 as py
from oracle import g2

from helpers import generated_base_multipliers, random_canonical_fr, scalars_from_ints

pytestmark = pytest.mark.gpu


def _dev(x):
    import torch
    if x.dtype == np.uint64:
        x = x.view(np.int64)
    return torch.from_numpy(x.copy()).cuda()


def _points_array(points):
    return np.frombuffer(b"".join(g2.g2_affine_bytes(p) for p in points), dtype=np.uint8).reshape(len(points), 200).copy()


def _image(p):
    return np.frombuffer(g2.g2_projective_bytes_normalised(p), dtype=np.uint64)


def _ints(scal):
    return [sum(int(v) << (64 * i) for i, v in enumerate(row)) for row in scal]


def test_generated_g2_bases_are_multiples_of_the_generator():
    from snarkvm_b200 import device
    n, seed = 64, 0xB200
    got = device.generate_bases_g2(n, seed).cpu().numpy()
    ks = generated_base_multipliers(seed, n)
    for i in (0, 1, 17, 63):
        assert g2.g2_affine_from_bytes(got[i].tobytes()) == g2.g2_mul(g2.G2_GEN, int(ks[i]))
    assert all(g2.g2_is_on_curve(g2.g2_affine_from_bytes(got[i].tobytes())) for i in range(n))


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 300, 2000])
def test_msm_g2_vs_oracle(n):
    """sizes on both sides of the reference's n < 32 special case; device-resident and host-buffer (FFI-shaped) entry points"""
    from snarkvm_b200 import cuda as shim, device
    from snarkvm_b200.algorithms import VariableBase
    bases = device.generate_bases_g2(n, seed=100 + n)
    bh = bases.cpu().numpy()
    pts = [g2.g2_affine_from_bytes(bh[i].tobytes()) for i in range(n)]
    scal = random_canonical_fr(n, seed=200 + n)
    want = _image(g2.standard_msm(pts, _ints(scal)))
    assert (device.msm_g2(bases, _dev(scal)) == want).all()
    assert (VariableBase.msm(bases, _dev(scal)) == want).all()            # dispatch by point type (200-byte rows)
    assert (shim.msm_g2(bh, scal) == want).all()
    assert (VariableBase.msm(bh, scal) == want).all()


def test_msm_g2_edge_cases():
    """scalars 0, 1, r − 1; points at infinity; a repeated point (doubling inside a bucket); P and −P with one scalar
    (cancellation to ∞); all scalars equal (one hot bucket per window); a sum that is ∞; fewer scalars than points"""
    from snarkvm_b200 import device
    n = 600
    bases = device.generate_bases_g2(n, seed=7).cpu().numpy()
    pts = [g2.g2_affine_from_bytes(bases[i].tobytes()) for i in range(n)]
    scal = random_canonical_fr(n, seed=8)
    scal[0:5] = 0
    scal[5:10] = scalars_from_ints([1])[0]
    scal[10:15] = scalars_from_ints([py.R_MOD - 1])[0]
    for i in range(20, 30):
        pts[i] = None
    for i in range(40, 80):
        pts[i] = pts[40]; scal[i] = scal[40]
    for i in range(100, 120, 2):
        pts[i + 1] = g2.g2_neg(pts[i]); scal[i + 1] = scal[i]
    arr = _points_array(pts)
    got = device.msm_g2(_dev(arr), _dev(scal))
    assert (got == _image(g2.standard_msm(pts, _ints(scal)))).all()
    same = np.tile(scal[300:301], (n, 1))
    assert (device.msm_g2(_dev(arr), _dev(same)) == _image(g2.standard_msm(pts, _ints(same)))).all()
    assert (device.msm_g2(_dev(arr), _dev(np.zeros((n, 4), dtype=np.uint64))) == _image(None)).all()
    pair = _points_array([pts[200], g2.g2_neg(pts[200])])
    assert (device.msm_g2(_dev(pair), _dev(np.tile(scal[200:201], (2, 1)))) == _image(None)).all()
    assert (device.msm_g2(_dev(arr), _dev(scal[:77])) == _image(g2.standard_msm(pts[:77], _ints(scal[:77])))).all()
    assert (device.msm_g2(_dev(arr[:0]), _dev(scal[:0])) == _image(None)).all()


def test_msm_g2_closed_form_2_16():
    """2^16 generated points P_i = h_i·G2 and uniform scalars: Σ s_i·P_i = (Σ s_i·h_i mod r)·G2 — one big-int dot product and one
    scalar multiplication check the whole device pipeline at a size the big-int Pippenger would take minutes for"""
    from snarkvm_b200 import device
    n, seed = 1 << 16, 4242
    bases = device.generate_bases_g2(n, seed)
    scal = random_canonical_fr(n, seed=99)
    ks = generated_base_multipliers(seed, n)
    dot = sum(int(k) * s for k, s in zip(ks, _ints(scal))) % py.R_MOD
    assert (device.msm_g2(bases, _dev(scal)) == _image(g2.g2_mul(g2.G2_GEN, dot))).all()


def test_msm_g2_rejects_bad_arguments():
    from snarkvm_b200 import CudaError, device
    bases = device.generate_bases_g2(8, seed=1)
    scal = random_canonical_fr(8, seed=2)
    with pytest.raises(ValueError):
        device.msm_g2(bases[:4], _dev(scal))                               # more scalars than points
    big = scal.copy(); big[3, 3] = np.uint64(1 << 62)                      # bit 254 set: not a canonical Fr
    with pytest.raises(CudaError):
        device.msm_g2(bases, _dev(big))

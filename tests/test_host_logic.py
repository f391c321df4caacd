"""Host-side logic that needs no GPU: domain sizing, shard ranges, the host fold of window sums, and the
world_size-2 (gloo) exchange that the NCCL path performs on the B200 box."""
import os
import random

import numpy as np
import pytest

from oracle import bls12_377 as py

from helpers import affine_array, oracle_bases, random_canonical_fr, scalars_from_ints


def test_evaluation_domain_new():
    from snarkvm_b200.algorithms import EvaluationDomain
    for num, size in ((0, 1), (1, 1), (2, 2), (3, 4), (4, 4), (5, 8), (1000, 1024), (1 << 20, 1 << 20), ((1 << 20) + 1, 1 << 21)):
        d = EvaluationDomain.new(num)
        assert d.size == size and d.log_size_of_group == size.bit_length() - 1
    assert EvaluationDomain.new((1 << 47) + 1) is None            # above TWO_ADICITY (domain.rs:125-127)
    assert EvaluationDomain.compute_size_of_domain(9) == 16


def test_shard_range_partitions():
    from snarkvm_b200.sharded import shard_range
    for n in (1, 7, 8, 1000, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) == (n + world - 1) // world


def _xyzz_affine_bytes(p):
    """XYZZ image (X, Y, ZZ, ZZZ) of an affine point (ZZ = ZZZ = 1) or infinity (all zero)."""
    if p is None:
        return b"\0" * 192
    one = py.fq_to_mont(1).to_bytes(48, "little")
    return py.fq_to_mont(p[0]).to_bytes(48, "little") + py.fq_to_mont(p[1]).to_bytes(48, "little") + one + one


def test_msm_finish_host_fold():
    """snarkvm_b200_msm_finish = Σ_w 2^{c·w}·S_w (batched.rs:404-413), pure host code — checked against Python."""
    from snarkvm_b200 import device
    rng = random.Random(3)
    for c, nwin in ((1, 3), (5, 4), (13, 20), (16, 16)):
        pts = [py.g1_mul(py.G1_GENERATOR, rng.randrange(1, 1 << 40)) for _ in range(nwin)]
        pts[1] = None
        sums = np.frombuffer(b"".join(_xyzz_affine_bytes(p) for p in pts), dtype=np.uint8).reshape(nwin, 192)
        expect = None
        for w, p in enumerate(pts):
            expect = py.g1_add(expect, py.g1_mul(p, 1 << (c * w)))
        assert device.msm_finish(sums, c).tobytes() == py.projective_bytes_normalised(expect)
    # c = 0: plain sum, including P + (−P) + P and the doubling branch
    p = py.g1_mul(py.G1_GENERATOR, 12345)
    for pts in ([p, py.g1_neg(p), p], [p, p], [None, None], [py.g1_neg(p), p]):
        sums = np.frombuffer(b"".join(_xyzz_affine_bytes(q) for q in pts), dtype=np.uint8).reshape(len(pts), 192)
        expect = None
        for q in pts:
            expect = py.g1_add(expect, q)
        assert device.msm_finish(sums, 0).tobytes() == py.projective_bytes_normalised(expect)


def _gloo_worker(rank, world, port, n, tmpdir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cpu
    from snarkvm_b200.sharded import combine_partials_host, shard_range
    cpu.set_num_threads(2)
    bases = oracle_bases(cpu, n, seed=11)
    scal = random_canonical_fr(n, seed=12)
    lo, hi = shard_range(n, rank, world)
    part = cpu.msm(bases[lo:hi], scal[lo:hi], 0)                 # this rank's shard, normalised (Z = R)
    x, y, z = (part[0:6], part[6:12], part[12:18])
    one = np.frombuffer(py.fq_to_mont(1).to_bytes(48, "little"), dtype=np.uint64)
    xyzz = np.zeros(24, dtype=np.uint64)
    if z.any():
        xyzz[0:6], xyzz[6:12], xyzz[12:18], xyzz[18:24] = x, y, one, one
    mine = torch.from_numpy(xyzz.view(np.int64).copy())
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)                              # the one exchange step of the sharded MSM
    partials = np.stack([g.numpy().view(np.uint64) for g in gathered])
    total = combine_partials_host(partials)
    full = cpu.msm(bases, scal, 0)
    np.save(os.path.join(tmpdir, f"ok_{rank}.npy"), np.array([int((total == full).all())]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_msm_exchange_gloo_world2(tmp_path, oracle_cpu):
    """N>1 path on CPU: 2 ranks shard the points, exchange their partial sums with one all-gather (gloo stands
    in for NCCL) and fold them with the product's host combine; result == unsharded oracle MSM."""
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, 200, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert np.load(tmp_path / f"ok_{r}.npy")[0] == 1


def test_kzg_argument_checks_need_no_gpu():
    """the reference's argument checks run before any device call (kzg10/mod.rs:134-137,166-169,283-290)"""
    import torch
    from snarkvm_b200 import algorithms as alg
    from snarkvm_b200.algorithms import KZG10, EvaluationDomain
    # host-side Montgomery helpers against the big-int root of trust
    rnd = random.Random(1)
    for _ in range(20):
        v = rnd.randrange(py.R_MOD)
        m = alg._fr_int_to_mont(v)
        assert sum(int(x) << (64 * i) for i, x in enumerate(m)) == py.fr_to_mont(v)
        assert alg._fr_mont_to_int(m) == v
    assert alg._R_MOD == py.R_MOD
    domain = EvaluationDomain.new(8)
    basis = torch.zeros(8 * 104, dtype=torch.uint8)
    evals = torch.zeros((8, 4), dtype=torch.int64)
    w = py.fr_root_of_unity(8)
    in_domain = alg._fr_int_to_mont(pow(w, 3, py.R_MOD))
    with pytest.raises(ValueError, match="Point cannot be in the domain"):
        KZG10.open_lagrange(basis, domain, evals, in_domain, alg._fr_int_to_mont(0))
    with pytest.raises(ValueError, match="must equal"):
        KZG10.open_lagrange(basis, domain, evals[:5], alg._fr_int_to_mont(5), alg._fr_int_to_mont(0))
    with pytest.raises(ValueError, match="Lagrange basis size"):
        KZG10.commit_lagrange(basis, evals[:3])                 # next_power_of_two(3) = 4 ≠ 8


def test_staging_copies_match_memcpy():
    """the pageable → pinned staging path of snarkvm_msm / snarkvm_ntt (copy-thread pool + non-temporal stores, csrc/hostcopy.cpp):
    contiguous and column-range copies of awkward sizes and alignments equal memcpy.  Host code only."""
    import ctypes
    from snarkvm_b200 import _lib
    bad = ctypes.c_uint32(99)
    _lib.check(_lib.lib().snarkvm_b200_selftest_host_copy(9 << 20, 0xC0DE, ctypes.byref(bad)))
    assert bad.value == 0


def test_msm_plan_rounds_the_size_to_the_nearest_power_of_two():
    """msm_make_plan (csrc/msm.cu): the window plan of a size just above 2^k is the plan of 2^k (a 2^20-coefficient polynomial plus a
    few blinding terms must not fall into the 2^21 plan), the switch happens at 2^(k + 1/2); batched.rs:390-393 picks c from
    ceil_log2 instead — parity is unaffected, any window size gives the same group element."""
    from snarkvm_b200 import device
    for k in (12, 16, 20, 22, 24):
        base = device.msm_plan(1 << k)
        assert device.msm_plan((1 << k) + 4) == {**base, "cap": device.msm_plan((1 << k) + 4)["cap"]}
        assert device.msm_plan(int(1.40 * (1 << k)))["c"] == base["c"]
        assert device.msm_plan(int(1.42 * (1 << k)))["c"] == device.msm_plan(1 << (k + 1))["c"]
    for n in (1, 2, 3, 7, 1000, (1 << 26)):
        p = device.msm_plan(n)
        assert p["nwin"] == 253 // p["c"] + 1 and p["c"] >= 2

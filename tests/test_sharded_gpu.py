"""N > 1 path on real GPUs: two ranks (NCCL) shard the points, all-gather their window sums over NVLink, add them
on the device and fold on the host; every rank must obtain the unsharded oracle result.  Skipped on 1-GPU boxes
(the driver's 1→8 scaling run and tests/test_host_logic.py's gloo test cover the logic there)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, tmpdir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from helpers import random_canonical_fr
    from snarkvm_b200 import device, sharded
    # the same global problem on every rank; each keeps only its shard in HBM
    bases = device.generate_bases(n, seed=77, device=f"cuda:{rank}")
    scal = random_canonical_fr(n, seed=78)
    lo, hi = sharded.shard_range(n, rank, world)
    dsc = torch.from_numpy(scal[lo:hi].view(np.int64).copy()).cuda(rank)
    got = sharded.msm_sharded(bases[lo:hi].contiguous(), dsc)
    if rank == 0:
        np.save(os.path.join(tmpdir, "bases.npy"), bases.cpu().numpy())
        np.save(os.path.join(tmpdir, "scal.npy"), scal)
    np.save(os.path.join(tmpdir, f"got_{rank}.npy"), got)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1 << 16, (1 << 16) + 1, 3])
def test_sharded_msm_two_gpus(tmp_path, oracle_cpu, n):
    """equal shards; ⌈n/2⌉ shards of different lengths whose own plans would differ (2^15 + 1 vs 2^15 points); and a 2 + 1 split"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, 29600 + os.getpid() % 1000, n, str(tmp_path)), nprocs=2, join=True)
    want = oracle_cpu.msm(np.load(tmp_path / "bases.npy"), np.load(tmp_path / "scal.npy"), 0)
    for r in range(2):
        assert (np.load(tmp_path / f"got_{r}.npy") == want).all(), r

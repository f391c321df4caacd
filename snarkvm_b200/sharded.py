"""Multi-GPU MSM: shard (bases, scalars) by contiguous point ranges, one process per GPU.

Replaces the reference's multi-GPU split (algorithms/cuda/cuda/snarkvm.cu:254-295: even split over
ngpus(), partial Jacobian points copied to the host and added there with point_t::dadd).  Here each
rank leaves its per-window XYZZ sums in HBM, the only exchange is one all-gather of nwin × 192 B per
rank over NCCL/NVLink, a device kernel adds the ranks window by window, and the ≤ 24 window sums are
folded on the host (Horner).  MSM is a sum, so no other collective exists on the path.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import device


def shard_range(npoints: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous ⌈n/G⌉ split, the reference's own (snarkvm.cu:254-269)."""
    per = (npoints + world - 1) // world
    lo = min(npoints, rank * per)
    return lo, min(npoints, lo + per)


def combine_partials_host(partials: np.ndarray) -> np.ndarray:
    """Σ of per-rank partial results given as XYZZ points [world, 192 B] on the host."""
    return device.msm_finish(partials, 0)


def msm_sharded(bases_shard: torch.Tensor, scalars_shard: torch.Tensor, group=None,
                stride: int = device.AFFINE_STRIDE) -> np.ndarray:
    """Every rank passes ITS shard (equal shard sizes ⇒ equal window plans); every rank returns the full sum."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    sums = device.msm_window_sums(bases_shard, scalars_shard, stride)
    npoints = (scalars_shard.numel() * scalars_shard.element_size()) // 32
    plan = device.msm_plan(npoints)
    if world > 1:
        gathered = torch.empty((world,) + tuple(sums.shape), dtype=sums.dtype, device=sums.device)
        dist.all_gather_into_tensor(gathered, sums, group=group)
        sums = device.xyzz_sum_ranks(gathered, world, plan["nwin"])
    return device.msm_finish(sums.cpu().numpy(), plan["c"])

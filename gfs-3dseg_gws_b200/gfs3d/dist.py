"""Multi-GPU plumbing (one process per GPU, torch.distributed).  The hot path shards by independent units
(SURVEY.md section 8e): point-cloud blocks for inference, points for k-means.  The only exchanges are
  * k-means: one all-reduce per Lloyd iteration of the packed fp64 [sums (K*D) | counts (K)] vector
  * timing:  max over ranks of a device-time measurement
Backend-agnostic (NCCL on the GPU box, gloo in the CPU tests)."""
from typing import Optional, Tuple

import torch


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous [lo, hi) slice of n_units owned by `rank`; sizes differ by at most one, earlier ranks get the extras"""
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_centroid_stats(sums: torch.Tensor, counts: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """sums (K, D) fp64, counts (K) int64 of this shard -> global sums / counts on every rank (one collective)"""
    import torch.distributed as dist
    K, D = sums.shape
    packed = torch.cat([sums.reshape(-1).double(), counts.double()])
    dist.all_reduce(packed, group=group)
    return packed[:K * D].reshape(K, D), packed[K * D:].round().long()


def max_over_ranks(value: float, device, group=None) -> float:
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])


def all_same(flag_local: bool, device, group=None) -> bool:
    """True iff flag_local is True on every rank (the labels-unchanged convergence test of the sharded Lloyd loop)"""
    import torch.distributed as dist
    t = torch.tensor([0.0 if flag_local else 1.0], dtype=torch.float64, device=device)
    dist.all_reduce(t, group=group)
    return float(t[0]) == 0.0

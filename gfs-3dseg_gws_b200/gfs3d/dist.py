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


class GradBucket:
    """ONE flat fp32 buffer that IS the gradient storage of all parameters (every p.grad is a view into it), so the
    data-parallel exchange of a training step (403 008 floats = 1.6 MB for the ScanNet model) is a single
    in-place all-reduce with no gather / scatter copies around it.  autograd accumulates into the views in place; call zero()
    instead of optimizer.zero_grad() (or zero_grad(set_to_none=False)).  BatchNorm statistics stay per replica (no SyncBN in
    the reference)."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero(self):
        self.flat.zero_()

    def views_intact(self) -> bool:
        """False if something replaced a .grad (e.g. zero_grad(set_to_none=True)): the bucket would then be stale"""
        lo, hi = self.flat.data_ptr(), self.flat.data_ptr() + self.flat.numel() * 4
        return all(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in self.params)

    def allreduce(self) -> int:
        """average over the ranks, in place; returns the number of floats exchanged (0 without an initialised process group)"""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return 0
        world = dist.get_world_size(self.group)
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)      # averaged inside the collective
        else:
            dist.all_reduce(self.flat, group=self.group)
            self.flat.div_(world)
        return self.flat.numel()


def allreduce_gradients(params, group=None):
    """data-parallel training: ONE all-reduce per step over a flat fp32 bucket of all gradients
    (398 144 floats = 1.6 MB for the S3DIS model: latency-bound on NVLink), averaged over the world.  BatchNorm statistics
    stay per replica (the reference has no SyncBN)."""
    import torch.distributed as dist
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not (dist.is_available() and dist.is_initialized()):
        return 0
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    dist.all_reduce(flat, group=group)
    flat.div_(dist.get_world_size(group))
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel()

"""GPU suite, N > 1: block-sharded inference and point-sharded k-means over NCCL (skipped on a 1-GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "gfs-3dseg_gws_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from gfs3d.dist import shard_range
    from gfs3d.kmeans import KMeans
    rs = np.random.RandomState(5)
    n, D, K = 40000, 192, 150
    cent = rs.randn(K, D).astype(np.float32)
    X = (cent[rs.randint(0, K, n)] + 0.35 * rs.randn(n, D)).astype(np.float32)
    init = X[rs.choice(n, K, replace=False)].copy()
    lo, hi = shard_range(n, rank, world)
    km = KMeans(n_clusters=K, init=init, shard=True).fit(X[lo:hi])
    ref = KMeans(n_clusters=K, init=init).fit(X) if rank == 0 else None
    ok = True
    if rank == 0:
        agree = float((ref.labels_[lo:hi] == km.labels_).mean())
        ok = agree >= 0.9999 and abs(ref.n_iter_ - km.n_iter_) <= 1 and np.abs(ref.cluster_centers_ - km.cluster_centers_).max() < 1e-4
        out["agree"] = agree
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_kmeans_matches_single_gpu():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, 29700 + os.getpid() % 200, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert all(out.get(r) for r in range(world)), dict(out)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_launches_follow_the_tensors_device_not_the_ambient_one():
    """a model / tensor on cuda:1 while the ambient device is cuda:0 (no torch.cuda.set_device): the C-ABI call must run on
    cuda:1's context and stream, and tensors on different devices must raise instead of launching"""
    for p in (ROOT, os.path.join(ROOT, "gfs-3dseg_gws_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from gfs3d import ops
    from oracle import gfs_oracle as O
    torch.cuda.set_device(0)
    x = O.synthetic_blocks(2, 256, seed=3)
    idx = ops.knn(x.to("cuda:1"), 20)
    assert idx.device == torch.device("cuda", 1)
    assert torch.equal(idx.cpu(), O.knn_exact(x, 20))
    with pytest.raises(RuntimeError, match="different devices"):
        ops.pointwise(x.to("cuda:1"), torch.randn(9, 128, device="cuda:0"), None)

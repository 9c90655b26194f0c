#!/usr/bin/env python
"""BASELINE.json configs[3]: global k-means (150 centroids, 192-d) over EdgeConv123-shaped features, per-GPU shard timing.

    python scripts/bench_kmeans.py [--n 500000] [--iters 10]          (one GPU: one shard of the 4 M / 8 split)
    torchrun --nproc-per-node N scripts/bench_kmeans.py --n 4000000   (points sharded over N GPUs, NCCL all-reduce)

Prints one JSON line: ms per Lloyd iteration (E-step + M-step + all-reduce), kernel times, achieved GB/s / TFLOP/s,
and sklearn's Lloyd iteration on a bounded sample of the host cores for comparison."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gfs-3dseg_gws_b200"))
from gfs3d import ops  # noqa: E402
from gfs3d.dist import allreduce_centroid_stats, shard_range  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=500000)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--cpu-sample", type=int, default=200000)
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
D, K = 192, 150
lo, hi = shard_range(a.n, rank, world)
n = (hi - lo + 3) // 4 * 4
g = torch.Generator(device=dev).manual_seed(99 + rank)
cent = torch.randn(K, D, device=dev, generator=g)
X = cent[torch.randint(0, K, (n,), device=dev, generator=g)] + 0.35 * torch.randn(n, D, device=dev, generator=g)
xt = X.t().contiguous()
centers = X[torch.randperm(n, device=dev, generator=g)[:K]].clone()
Kp = (K + 3) // 4 * 4


def lloyd(centers):
    ct = torch.zeros(D, Kp, device=dev)
    ct[:, :K] = centers.t()
    labels = ops.kmeans_assign(xt, ct, K)
    sums, counts = ops.kmeans_accumulate(X, labels, K)
    if world > 1:
        sums, counts = allreduce_centroid_stats(sums, counts)
    return (sums / counts.clamp_min(1)[:, None]).float(), labels


for _ in range(3):
    centers, _ = lloyd(centers)
torch.cuda.synchronize()
ops.PROFILE = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    centers, labels = lloyd(centers)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
prof = {k: sum(s.elapsed_time(e) for s, e in v) / a.iters for k, v in ops.PROFILE.items()}
t = torch.tensor([ms], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    am = prof.get("gfs_kmeans_assign_tc", prof.get("gfs_kmeans_assign"))      # tensor-core E-step when the shape is eligible
    cm = prof["gfs_kmeans_accumulate"]
    line = {"metric": "kmeans_lloyd_iteration_ms", "value": float(t[0]), "unit": "ms", "n_gpus": world, "points_total": a.n,
            "points_per_gpu": n, "centroids": K, "dim": D, "higher_is_better": False,
            "assign": {"ms": am, "fp32_tflops": 2.0 * n * D * 192 / am / 1e9, "algorithmic_gbs": (n * D * 4 + n * 4) / am / 1e6},
            "accumulate": {"ms": cm, "algorithmic_gbs": (n * D * 4 + n * 4) / cm / 1e6, "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "allreduce_bytes": (K * D + K) * 8 if world > 1 else 0}
    # k-means++ seeding: one pass over the resident channel-major copy per centre (gfs_kmeans_pp_trial)
    from gfs3d.kmeans import KMeans as GKMeans
    seeder = GKMeans(n_clusters=K, init="k-means++", random_state=0)
    seeder._seed_plusplus(X[:40000], xt[:, :40000].contiguous(), np.random.RandomState(0))      # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    seeder._seed_plusplus(X, xt, np.random.RandomState(0))
    torch.cuda.synchronize()
    seed_s = time.perf_counter() - t0
    line["seeding"] = {"seconds": seed_s, "ms_per_centre": 1e3 * seed_s / K, "trials_per_centre": 2 + int(np.log(K)),
                       "algorithmic_gbs": K * n * D * 4 / 1e9 / seed_s, "hbm_peak_gbs": peaks.get("hbm_gbs")}
    if a.cpu_sample:
        from sklearn.cluster import KMeans
        rs = np.random.RandomState(0)
        c = rs.randn(K, D).astype(np.float32)
        Xc = (c[rs.randint(0, K, a.cpu_sample)] + 0.35 * rs.randn(a.cpu_sample, D)).astype(np.float32)
        init = Xc[rs.choice(a.cpu_sample, K, replace=False)]
        t0 = time.perf_counter()
        km = KMeans(n_clusters=K, init=init, n_init=1, max_iter=5, tol=0).fit(Xc)
        dt = time.perf_counter() - t0
        from sklearn.cluster import kmeans_plusplus
        t0 = time.perf_counter()
        kmeans_plusplus(Xc, K, random_state=np.random.RandomState(0))
        line["seeding"]["cpu_sklearn_seconds_scaled_to_points_per_gpu"] = (time.perf_counter() - t0) * n / a.cpu_sample
        line["cpu_sklearn"] = {"ms_per_iter": 1e3 * dt / max(1, km.n_iter_), "points": a.cpu_sample, "cores": os.cpu_count(),
                               "ms_per_iter_scaled_to_points_per_gpu": 1e3 * dt / max(1, km.n_iter_) * n / a.cpu_sample}
    print(json.dumps(line))
if world > 1:
    dist.destroy_process_group()

"""KMeans -- drop-in for `sklearn.cluster.KMeans` as get_basis.py:210 uses it: `KMeans(n_clusters=, init='k-means++').fit(X).labels_`.

Lloyd iterations run on the GPU: gfs_kmeans_assign (fp32 pinned-order E-step) + gfs_kmeans_accumulate (deterministic
sums) per iteration; the control flow mirrors sklearn 1.9.0 (_kmeans.py:1487-1490 mean-centring, :289-296 tolerance,
:630-759 Lloyd loop / convergence, _k_means_common.pyx:167-260 empty-cluster relocation and centre averaging).

Multi-GPU: pass `process_group` (or have torch.distributed initialised and shard=True): X is then THIS rank's shard of
the points; the only exchange per iteration is an NCCL all-reduce of the (K*D + K) fp64 sums/counts (+ the mean/variance
once), every rank applies the identical centroid update.

Inject for the reference script without editing it:   import get_basis; get_basis.KMeans = gfs3d.kmeans.KMeans
"""
from typing import Optional

import numpy as np
import torch

from . import ops
from .dist import all_same, allreduce_centroid_stats


def _dist():
    import torch.distributed as dist
    return dist


class KMeans:
    def __init__(self, n_clusters=8, *, init="k-means++", n_init="auto", max_iter=300, tol=1e-4, verbose=0,
                 random_state=None, copy_x=True, algorithm="lloyd", device=None, shard=False, process_group=None):
        if algorithm not in ("lloyd", "auto", "full"):
            raise NotImplementedError(f"algorithm={algorithm!r}: only Lloyd is built")
        self.n_clusters, self.init, self.n_init = n_clusters, init, n_init
        self.max_iter, self.tol, self.verbose, self.random_state = max_iter, tol, verbose, random_state
        self.device = device
        self.shard = shard or process_group is not None
        self.process_group = process_group

    # ------------------------------------------------------------------ helpers
    def _allreduce(self, t):
        if self.shard:
            _dist().all_reduce(t, group=self.process_group)
        return t

    def _rng(self):
        rs = self.random_state
        if rs is None:
            return np.random.mtrand._rand          # numpy global state, as sklearn's check_random_state(None)
        if isinstance(rs, np.random.RandomState):
            return rs
        return np.random.RandomState(rs)

    def _seed_plusplus(self, Xc, xt, rng):
        """k-means++ (sklearn _kmeans.py:_kmeans_plusplus): per centre, 2 + ln K candidates drawn with probability
        proportional to the current squared distance (host RNG consumed in sklearn's order, cumulative sum and search on the
        device), all candidates scored in ONE pass over the resident channel-major copy (gfs_kmeans_pp_trial), the best kept."""
        n, D = Xc.shape
        K = self.n_clusters
        trials = 2 + int(np.log(K))
        if trials > 8:
            raise NotImplementedError(f"k-means++ with {trials} local trials (n_clusters={K}) is not built")
        dev = Xc.device
        npad = xt.shape[1]
        xsq = torch.zeros(npad, dtype=torch.float32, device=dev)
        xsq[:n] = (Xc * Xc).sum(1)
        centers = torch.empty(K, D, dtype=torch.float32, device=dev)
        m = [torch.empty(8, npad, dtype=torch.float32, device=dev) for _ in range(2)]      # double buffer: closest lives in one
        pots = torch.zeros(8, dtype=torch.float64, device=dev)
        first = int(rng.choice(n))
        centers[0] = Xc[first]
        ops.kmeans_pp_trial(xt, n, xsq, centers[0:1].contiguous(), None, m[0], pots)
        closest, pot, cur = m[0][0], pots[0].clone(), 0
        for c in range(1, K):
            rv = torch.from_numpy(rng.uniform(size=trials)).to(dev) * pot
            cand = torch.searchsorted(torch.cumsum(closest[:n].double(), 0), rv).clamp_max_(n - 1)
            pots.zero_()
            ops.kmeans_pp_trial(xt, n, xsq, Xc[cand].contiguous(), closest, m[cur ^ 1], pots)
            best = int(torch.argmin(pots[:trials]))
            cur ^= 1
            pot, closest = pots[best].clone(), m[cur][best]
            centers[c] = Xc[cand[best]]
        return centers

    # ------------------------------------------------------------------ fit
    def fit(self, X, y=None, sample_weight=None):
        if sample_weight is not None:
            raise NotImplementedError("sample_weight is not built (get_basis.py never passes it)")
        dev = torch.device(self.device) if self.device is not None else torch.device("cuda", torch.cuda.current_device())
        if isinstance(X, np.ndarray):
            Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).to(dev, non_blocking=True)
        else:
            Xd = X.to(dev, torch.float32).contiguous()
        n, D = Xd.shape
        K = self.n_clusters
        if K > 192:
            raise NotImplementedError(f"n_clusters={K} > 192 is not built")

        # mean-centring and tolerance (global over all shards)
        stat = torch.cat([Xd.double().sum(0), (Xd.double() ** 2).sum(0), torch.tensor([float(n)], dtype=torch.float64, device=dev)])
        self._allreduce(stat)
        n_tot = float(stat[-1])
        mean = (stat[:D] / n_tot)
        var = stat[D:2 * D] / n_tot - mean ** 2
        tol_abs = float(var.mean()) * self.tol
        mean32 = mean.float()
        Xc = Xd - mean32
        npad = (n + 3) // 4 * 4
        xt = torch.zeros(D, npad, dtype=torch.float32, device=dev)
        xt[:, :n] = Xc.t()

        # initial centres
        if isinstance(self.init, str) and self.init == "k-means++":
            if self.shard and _dist().get_rank(self.process_group) != 0:
                centers = torch.empty(K, D, dtype=torch.float32, device=dev)
            else:
                centers = self._seed_plusplus(Xc, xt, self._rng())  # multi-GPU: seeded from rank 0's shard, then broadcast
            if self.shard:
                _dist().broadcast(centers, src=_dist().get_global_rank(self.process_group, 0) if self.process_group else 0,
                                  group=self.process_group)
        elif isinstance(self.init, (np.ndarray, torch.Tensor)):
            centers = torch.as_tensor(self.init, dtype=torch.float32).to(dev) - mean32
            if centers.shape != (K, D):
                raise ValueError(f"init has shape {tuple(centers.shape)}, expected {(K, D)}")
        else:
            raise NotImplementedError(f"init={self.init!r} is not built")

        Kp = (K + 3) // 4 * 4
        labels_old = torch.full((npad,), -1, dtype=torch.int32, device=dev)
        strict = False
        n_iter = 0
        for n_iter in range(1, self.max_iter + 1):
            ct = torch.zeros(D, Kp, dtype=torch.float32, device=dev)
            ct[:, :K] = centers.t()
            labels = ops.kmeans_assign(xt, ct, K)
            sums, counts = ops.kmeans_accumulate(Xc, labels, K, n_valid=n)
            if self.shard:
                sums, counts = allreduce_centroid_stats(sums, counts, self.process_group)
            if bool((counts == 0).any()):
                sums, counts = self._relocate_empty(Xc, n, labels, centers, sums, counts)
            new = torch.where(counts[:, None] > 0, sums / counts.clamp_min(1)[:, None].double(), torch.zeros_like(sums)).float()
            if bool((counts <= 0).any()):
                new[counts <= 0] = new[int(torch.argmax(counts))]
            shift = float(((new - centers).double() ** 2).sum())
            same = torch.equal(labels[:n], labels_old[:n])
            if self.shard:
                same = all_same(same, dev, self.process_group)
            centers = new
            if self.verbose:
                print(f"Iteration {n_iter - 1}, center shift {shift:.6g}")
            if same:
                strict = True
                break
            if shift <= tol_abs:
                break
            labels_old = labels
        if not strict:
            ct = torch.zeros(D, Kp, dtype=torch.float32, device=dev)
            ct[:, :K] = centers.t()
            labels = ops.kmeans_assign(xt, ct, K)
        self.labels_ = labels[:n].cpu().numpy().astype(np.int32)
        self.labels_device_ = labels[:n]
        self.cluster_centers_ = (centers + mean32).cpu().numpy()
        self.n_iter_ = n_iter
        self.n_features_in_ = D
        return self

    def _relocate_empty(self, Xc, n, labels, centers, sums, counts):
        """sklearn _k_means_common.pyx:167-211: every empty cluster is moved onto one of the points farthest from its own
        (old) centre, and that point is taken out of its old cluster's sum.  Rare path, plain device ops.  Sharded: each
        rank offers its n_empty farthest points (distance, old label, coordinates), the candidates are all-gathered and
        every rank applies the same global choice to the (already all-reduced) sums / counts."""
        empty = (counts == 0).nonzero().flatten()
        ne = int(empty.numel())
        D = Xc.shape[1]
        dist = ((Xc - centers[labels[:n].long()]) ** 2).sum(1)
        m = min(ne, n)
        vals, idxs = torch.topk(dist, m)
        cand = torch.full((ne, 2 + D), float("-inf"), dtype=torch.float64, device=Xc.device)
        cand[:m, 0] = vals.double()
        cand[:m, 1] = labels[idxs].double()
        cand[:m, 2:] = Xc[idxs].double()
        if self.shard:
            world = _dist().get_world_size(self.process_group)
            gathered = [torch.empty_like(cand) for _ in range(world)]
            _dist().all_gather(gathered, cand, group=self.process_group)
            cand = torch.cat(gathered, 0)
        if float(cand[:, 0].max()) <= 0:
            return sums, counts
        order = torch.argsort(cand[:, 0], descending=True, stable=True)[:ne]
        for e, row in zip(empty.tolist(), order.tolist()):
            old = int(cand[row, 1])
            x = cand[row, 2:]
            sums[old] -= x
            sums[e] = x
            counts[e] = 1
            counts[old] -= 1
        return sums, counts

    def fit_predict(self, X, y=None, sample_weight=None):
        return self.fit(X, sample_weight=sample_weight).labels_

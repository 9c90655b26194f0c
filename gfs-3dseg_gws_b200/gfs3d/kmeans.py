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

    def _world(self):
        if not self.shard:
            return 1, 0
        d = _dist()
        return d.get_world_size(self.process_group), d.get_rank(self.process_group)

    def _seed_plusplus(self, Xc, xt, rng, n=None):
        """k-means++ exactly as sklearn's _kmeans_plusplus consumes its random stream (sklearn 1.9.0 _kmeans.py:_kmeans_plusplus):
        first centre through `choice(n, p=w/sum(w))`, then `uniform(size=2+ln K)` per centre; candidates = searchsorted of
        u * potential in the cumulative sum of the closest squared distances; all candidates of a centre scored in ONE pass over
        the resident channel-major copy (gfs_kmeans_pp_trial: fp64-accumulated distances rounded to fp32, like sklearn's upcast
        path), the best kept.  What is NOT bit-pinned: sklearn sums the potential with a float32 BLAS dot and the cumulative sum in
        float32; here both are fp64 (parallel, order-independent to 1e-16), so a pick can differ where u * pot lands within that
        rounding of a boundary (tests/golden: 31 of 32 seeded problems give identical picks for all centres).

        Sharded (multi-GPU): the points of all ranks form ONE sequence in rank order; every rank searches its own slice of the
        global cumulative sum, the owner contributes the candidate rows (all-reduce), potentials are all-reduced: the picks equal
        the single-GPU picks.  The random numbers are drawn on rank 0 and broadcast.  No host synchronisation inside the loop."""
        D = Xc.shape[1]
        n = Xc.shape[0] if n is None else n
        K = self.n_clusters
        trials = 2 + int(np.log(K))
        if trials > 8:
            raise NotImplementedError(f"k-means++ with {trials} local trials (n_clusters={K}) is not built")
        dev = Xc.device
        world, rank = self._world()
        sizes = torch.tensor([n], dtype=torch.int64, device=dev)
        if world > 1:
            allsz = [torch.empty_like(sizes) for _ in range(world)]
            _dist().all_gather(allsz, sizes, group=self.process_group)
            sizes = torch.cat(allsz)
        sizes = sizes.tolist()
        n_tot, lo = sum(sizes), sum(sizes[:rank])
        # the random stream (rank 0's generator; the other ranks receive the draws)
        draws = torch.empty(1 + (K - 1) * trials, dtype=torch.float64)
        if rank == 0:
            w = np.ones(n_tot, dtype=np.float32)
            draws[0] = float(rng.choice(n_tot, p=w / w.sum()))
            draws[1:] = torch.from_numpy(rng.uniform(size=(K - 1) * trials)) if K > 1 else draws[1:]
        draws = draws.to(dev)
        if world > 1:
            _dist().broadcast(draws, src=_dist().get_global_rank(self.process_group, 0) if self.process_group else 0,
                              group=self.process_group)
        U = draws[1:].view(K - 1, trials)
        npad = xt.shape[1]
        xsq = torch.zeros(npad, dtype=torch.float64, device=dev)
        xsq[:n] = (Xc[:n].double() ** 2).sum(1)
        centers = torch.empty(K, D, dtype=torch.float32, device=dev)
        m = torch.empty(8, npad, dtype=torch.float32, device=dev)
        pots = torch.zeros(8, dtype=torch.float64, device=dev)

        def owned_rows(pos, mine):
            """(T, D) rows X[pos] where this rank owns the pick, zero elsewhere; summed over the ranks"""
            rows = Xc[pos.clamp(0, n - 1)] * mine.to(torch.float32).unsqueeze(1)
            if world > 1:
                _dist().all_reduce(rows, group=self.process_group)
            return rows.contiguous()

        first = draws[0].long() - lo
        centers[0] = owned_rows(first.view(1), ((first >= 0) & (first < n)).view(1))[0]
        ops.kmeans_pp_trial(xt, n, xsq, centers[0:1].contiguous(), None, m, pots)
        if world > 1:
            _dist().all_reduce(pots, group=self.process_group)
        closest, pot = m[0].clone(), pots[0].clone()
        for c in range(1, K):
            cs = torch.cumsum(closest[:n].double(), 0)             # this rank's slice of the global cumulative sum, minus P[rank]
            tots = cs[-1:].contiguous()
            if world > 1:
                gathered = [torch.empty_like(tots) for _ in range(world)]
                _dist().all_gather(gathered, tots, group=self.process_group)
                tots = torch.cat(gathered)
            P = torch.cumsum(tots, 0)                              # identical on every rank: ownership is decided from the same numbers
            rv = U[c - 1] * pot
            owner = torch.searchsorted(P, rv).clamp_max(world - 1)  # first rank whose slice reaches rv (the last one if none: numpy's clip)
            mine = owner == rank
            pos = torch.searchsorted(cs, rv - (P[rank] - tots[rank])).clamp_max(n - 1)   # first i with cumsum[i] >= rv (side='left')
            cand = owned_rows(pos, mine)
            pots.zero_()
            ops.kmeans_pp_trial(xt, n, xsq, cand, closest, m, pots)
            if world > 1:
                _dist().all_reduce(pots, group=self.process_group)
            best = torch.argmin(pots[:trials]).view(1)
            pot = pots.index_select(0, best)[0]
            closest = m.index_select(0, best)[0]
            centers[c] = cand.index_select(0, best)[0]
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
            centers = self._seed_plusplus(Xc, xt, self._rng(), n)
        elif isinstance(self.init, (np.ndarray, torch.Tensor)):
            centers = torch.as_tensor(self.init, dtype=torch.float32).to(dev) - mean32
            if centers.shape != (K, D):
                raise ValueError(f"init has shape {tuple(centers.shape)}, expected {(K, D)}")
        else:
            raise NotImplementedError(f"init={self.init!r} is not built")

        # Lloyd iterations.  Per iteration: E-step (4 launches), M-step sums (2), pack (1), ONE collective of the packed
        # [sums | counts | labels changed] buffer when sharded, centre update (1) and ONE device->host read of
        # (labels changed, centre shift, empty clusters).  All buffers are allocated once.
        Kp = (K + 3) // 4 * 4
        ct = torch.zeros(D, Kp, dtype=torch.float32, device=dev)
        ct[:, :K] = centers.t()
        centers = centers.contiguous()
        centers_new = torch.empty_like(centers)
        labels_old = torch.full((npad,), -1, dtype=torch.int32, device=dev)
        packed = torch.empty(K * D + K + 1, dtype=torch.float64, device=dev)
        sums = packed[:K * D].view(K, D)
        counts = torch.empty(K, dtype=torch.int64, device=dev)
        acc_ws = ops.kmeans_accumulate_workspace(K, D, dev)
        scratch = torch.zeros(2, dtype=torch.int64, device=dev)
        result = torch.empty(3, dtype=torch.float64, device=dev)
        # E-step over the channel-major copy: tensor-core product when the shape allows, same labels as the fp32 kernel.  (The
        # row-major entry, ops.kmeans_assign_rows, measured slower: 32 rows per load instruction instead of one 128-byte line.)
        assign = lambda: ops.kmeans_assign(xt, ct, K)
        strict = False
        n_iter = 0
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for n_iter in range(1, self.max_iter + 1):
            labels = assign()                                          # ct holds the current centres, transposed
            ops.kmeans_accumulate(Xc, labels, K, n_valid=n, sums=sums, counts=counts, ws=acc_ws)
            ops.kmeans_pack(labels, labels_old, n, counts, K, D, packed, scratch)
            if self.shard:
                _dist().all_reduce(packed, group=self.process_group)
            ops.kmeans_update(packed, centers, K, D, Kp, centers_new, ct, result)
            n_changed, shift_h, n_empty = result.tolist()
            if n_empty > 0:           # rare: sklearn moves the empty clusters onto the points farthest from their centres
                gsums = packed[:K * D].view(K, D).clone()
                gcounts = packed[K * D:K * D + K].round().long()
                gsums, gcounts = self._relocate_empty(Xc, n, labels, centers, gsums, gcounts)
                # clusters that are still empty keep the zero sum (sklearn _k_means_common.pyx:_average_centers)
                new = torch.where(gcounts[:, None] > 0, gsums / gcounts.clamp_min(1)[:, None].double(), torch.zeros_like(gsums)).float()
                shift_h = float(((new - centers).double() ** 2).sum())
                centers_new.copy_(new)
                ct[:, :K] = new.t()
            centers, centers_new = centers_new, centers
            if self.verbose:
                print(f"Iteration {n_iter - 1}, center shift {shift_h:.6g}")
            if n_changed == 0:
                strict = True
                break
            if shift_h <= tol_abs:
                break
            labels_old = labels
        ev1.record()
        if not strict:
            labels = assign()                                          # ct already holds the final centres
        self.labels_ = labels[:n].cpu().numpy().astype(np.int32)
        self.lloyd_ms_ = ev0.elapsed_time(ev1)          # device time of the Lloyd loop (the .cpu() above synchronised)
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

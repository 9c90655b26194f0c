"""GW basis construction -- the compute of get_basis.py:112-222 (`Get_GlobalProto_GlobalKmeans`) as a library call.

    features  = EdgeConv123 of every base-class point     (get_basis.py:162-183; here kept on the device, row N3 of SURVEY 8f)
    subsample = <= max_num points per class               (get_basis.py:189-198)
    labels    = KMeans(n_clusters=num_cnt, init='k-means++').fit(point_feat).labels_   (get_basis.py:210-212, GPU Lloyd)
    protos    = per-cluster mean                          (get_basis.py:27-44  Kmean2Proto)
    basis     = rank-truncated SVD reconstruction         (get_basis.py:50-71  compute_svd; numpy, parity-critical cut-off)

The data loading / argparse / pickling of get_basis.py stay with the reference script (out of scope).
"""
from typing import Iterable, Optional, Tuple

import numpy as np
import torch

from .kmeans import KMeans


def kmean_to_proto(feat: np.ndarray, labels: np.ndarray, num_cnt: int) -> np.ndarray:
    """(N, D) features + (N,) cluster labels -> (num_cnt, D) cluster means; an empty cluster is an error (get_basis.py:37)"""
    protos = np.empty((num_cnt, feat.shape[1]), dtype=feat.dtype)
    for c in range(num_cnt):
        members = labels == c
        if not members.any():
            raise AssertionError(f"k-means cluster {c} is empty")
        protos[c] = feat[members].mean(axis=0)
    return protos


def svd_reconstruct(protos: np.ndarray, energy: float = 0.95) -> np.ndarray:
    """keep the leading singular directions holding > `energy` of the singular-value mass and reconstruct -> (num_cnt, D)"""
    u, s, vh = np.linalg.svd(protos.T, full_matrices=False)
    mass = np.cumsum(s)
    r = int(np.argmax(mass > energy * s.sum())) if (mass > energy * s.sum()).any() else len(s) - 1
    return (u[:, : r + 1] @ np.diag(s[: r + 1]) @ vh[: r + 1, :]).T


@torch.no_grad()
def collect_edgeconv_features(encoder, blocks: Iterable[Tuple[torch.Tensor, torch.Tensor]], num_classes: int,
                              max_num: int = 300000, rng: Optional[np.random.RandomState] = None) -> torch.Tensor:
    """encoder: model.dgcnn.DGCNN(return_edgeconvs=True) in eval mode.  blocks yields (points (B,C,N), labels (B,N)).
    Returns the (n, 192) fp32 device matrix of per-class-subsampled EdgeConv123 features, classes 1..num_classes-1 in order
    (class 0 = background is skipped, get_basis.py:176-178)."""
    rng = rng or np.random
    per_class = {c: [] for c in range(1, num_classes)}
    for pts, lab in blocks:
        ecs, _ = encoder(pts.cuda(non_blocking=True))
        feat = torch.cat(ecs, dim=1).permute(0, 2, 1).reshape(-1, 64 * len(ecs))        # (B*N, 192)
        lab = lab.cuda(non_blocking=True).reshape(-1)
        for c in per_class:
            m = lab == c
            if bool(m.any()):
                per_class[c].append(feat[m])
    out = []
    for c in sorted(per_class):
        if not per_class[c]:
            continue
        f = torch.cat(per_class[c], dim=0)
        if f.shape[0] > max_num:
            keep = rng.choice(np.arange(f.shape[0]), max_num, replace=False)
            f = f[torch.from_numpy(np.sort(keep)).to(f.device)]
        out.append(f)
    return torch.cat(out, dim=0).contiguous()


def build_gw_basis(point_feat, num_cnt: int, energy: float = 0.95, init="k-means++", random_state=None, shard: bool = False):
    """point_feat (n, D) device tensor or ndarray -> (basis (num_cnt, D) fp32 ndarray, fitted KMeans)"""
    km = KMeans(n_clusters=num_cnt, init=init, random_state=random_state, shard=shard).fit(point_feat)
    feat_np = point_feat.detach().cpu().numpy() if isinstance(point_feat, torch.Tensor) else np.asarray(point_feat)
    if shard:
        raise NotImplementedError("sharded proto averaging is not built: gather labels or run Kmean2Proto per rank + all-reduce")
    protos = kmean_to_proto(feat_np, km.labels_, num_cnt)
    return svd_reconstruct(protos, energy).astype(np.float32), km

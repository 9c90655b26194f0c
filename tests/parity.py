"""Parity metrics shared by the CPU and GPU suites (near-tie classification)."""
import numpy as np
import torch


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max-norm relative error  max|a-b| / max|b|  (the tolerance quoted in BASELINE.json north_star)."""
    a = a.double().cpu()
    b = b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def knn_set_mismatch(idx_a, idx_b):
    """rows whose neighbour SETS differ; idx: (B,N,k) integer arrays."""
    a = np.sort(np.asarray(idx_a, dtype=np.int64), axis=-1)
    b = np.sort(np.asarray(idx_b, dtype=np.int64), axis=-1)
    return np.argwhere((a != b).any(axis=-1))


def knn_classify_mismatches(x: torch.Tensor, idx_a, idx_b, k: int):
    """For every row whose neighbour set differs, decide whether the disagreement is explained by fp32 rounding
    of the distance formula (a near-tie at the k-th/(k+1)-th boundary) or is a real error.

    x: (B,C,N) fp32.  Distances are re-evaluated in fp64; a row is 'near-tie' when every index that is in one set
    but not the other has an fp64 distance within `band` of the fp64 k-th distance, where
    band = 8 * eps32 * (|xx_i| + |xx_j| + 2|x_i.x_j|)  (the rounding budget of -xx_i + 2 x_i.x_j - xx_j).
    Returns (n_rows_mismatch, n_near_tie, n_real)."""
    rows = knn_set_mismatch(idx_a, idx_b)
    xd = x.double()
    eps = float(np.finfo(np.float32).eps)
    near = real = 0
    for b, i in rows:
        xi = xd[b, :, i]
        dots = xi @ xd[b]
        xx = (xd[b] ** 2).sum(0)
        d = -xx[i] + 2 * dots - xx
        kth = torch.topk(d, k + 1).values
        band = 8 * eps * (xx[i] + xx + 2 * dots.abs())
        sa, sb = set(np.asarray(idx_a)[b, i].tolist()), set(np.asarray(idx_b)[b, i].tolist())
        ok = True
        for j in sa ^ sb:
            if abs(float(d[j] - kth[k - 1])) > float(band[j]) and abs(float(d[j] - kth[k])) > float(band[j]):
                ok = False
        near += ok
        real += (not ok)
    return len(rows), near, real


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """relative Frobenius error ||a-b|| / ||b|| -- for gradients, where an fp32-vs-fp64 arg-max near-tie legitimately re-routes
    a single element's gradient (a large max-norm difference at isolated elements, negligible in norm)"""
    a = a.double().cpu()
    b = b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

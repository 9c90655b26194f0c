"""GPU suite: k-means E/M kernels vs the pinned-order C oracle (labels bit-exact given identical centroids)."""
import numpy as np
import pytest
import torch

from oracle import gfs_oracle as O

pytestmark = pytest.mark.gpu


def _mixture(n, D, K, seed):
    rs = np.random.RandomState(seed)
    cent = rs.randn(K, D).astype(np.float32)
    lab = rs.randint(0, K, size=n)
    return (cent[lab] + 0.35 * rs.randn(n, D)).astype(np.float32), rs


@pytest.mark.parametrize("n,D,K", [(6000, 192, 150), (20000, 192, 180), (4096, 64, 20), (1000, 192, 150), (64, 192, 7)])
def test_assign_bit_exact_and_sums(n, D, K):
    from gfs3d import ops
    X, rs = _mixture(n, D, K, seed=n + K)
    C = X[rs.choice(n, K, replace=False)].copy()
    ref_labels, ref_score = O.kmeans_assign_exact(X, C, return_score=True)
    Kp = (K + 3) // 4 * 4
    ct = np.zeros((D, Kp), np.float32)
    ct[:, :K] = C.T
    xt = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
    labels, score = ops.kmeans_assign(xt, torch.from_numpy(ct).cuda(), K, want_score=True)
    assert np.array_equal(labels.cpu().numpy(), ref_labels), "labels must be bit-exact given identical fp32 centroids"
    assert np.array_equal(score.cpu().numpy(), ref_score)
    sums, counts = ops.kmeans_accumulate(torch.from_numpy(X).cuda(), labels, K)
    rs_, rc_ = O.kmeans_accumulate_exact(X, ref_labels, K)
    assert np.array_equal(counts.cpu().numpy(), rc_)
    assert np.abs(sums.cpu().numpy() - rs_).max() <= 1e-3 * max(1.0, np.abs(rs_).max())
    # deterministic: same bits on a second run
    sums2, _ = ops.kmeans_accumulate(torch.from_numpy(X).cuda(), labels, K)
    assert torch.equal(sums, sums2)


def test_assign_ties_lowest_index():
    from gfs3d import ops
    xt = torch.zeros(8, 64, device="cuda")
    ct = torch.zeros(8, 8, device="cuda")
    assert int(ops.kmeans_assign(xt, ct, 5).max()) == 0

"""GPU suite at BASELINE.json's full sizes, through size-independent properties (the oracle only sees bounded samples)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import gfs_oracle as O
from parity import rel_err

pytestmark = pytest.mark.gpu


def _model(golden_sd, gp):
    from model.capl import mpti_net_Point_GeoAsWeight_v2
    args = SimpleNamespace(edgeconv_widths=[[64, 64]] * 3, dgcnn_mlp_widths=[512, 256], pc_in_dim=9, dgcnn_k=20,
                           base_widths=[128, 64], output_dim=64, eval_weight=1.2)
    m = mpti_net_Point_GeoAsWeight_v2(classes=13, criterion=torch.nn.CrossEntropyLoss(ignore_index=255), args=args, base_num=7,
                                      gp=gp.cuda(), energy=0.9)
    m.load_state_dict(golden_sd("gfs_s3dis_weights"))
    return m.cuda().eval()


def test_config2_full_batch_properties(golden_sd):
    """configs[1]: B = 32 blocks x 2048 points.  (a) kNN structure on every row, bit-exact vs the oracle on two blocks;
    (b) blocks are independent: a block's logits do not depend on what else is in the batch (bit-identical);
    (c) label agreement vs the oracle on a sample of blocks."""
    from gfs3d import ops
    from gfs3d.synthetic import synthetic_blocks
    B, N, k = 32, 2048, 20
    x = synthetic_blocks(B, N, seed=4242)
    xc = x.cuda()
    idx, dist = ops.knn(xc, k, return_dist=True)
    ii, dd = idx.cpu().numpy(), dist.cpu().numpy()
    assert ii.min() >= 0 and ii.max() < N
    assert (dd[..., :-1] >= dd[..., 1:]).all(), "neighbours must be sorted nearest first"
    assert (ii == np.arange(N)[None, :, None]).any(-1).all(), "every point is its own neighbour (distance 0)"
    assert (np.sort(ii, -1)[..., 1:] != np.sort(ii, -1)[..., :-1]).all(), "no index may appear twice in a row"
    for b in (0, 31):
        ref_i, ref_d = O.knn_exact(x[b:b + 1], k, return_dist=True)
        assert np.array_equal(ii[b], ref_i[0].numpy()) and np.array_equal(dd[b], ref_d[0].numpy())

    g = torch.Generator().manual_seed(11)
    gp = torch.randn(150, 192, generator=torch.Generator().manual_seed(7))
    gened = torch.nn.functional.normalize(torch.randn(13, 128, generator=g), dim=1)
    coding = (torch.rand(13, 150, generator=g) < 0.3).float()
    m = _model(golden_sd, gp)
    kw = dict(y=None, eval_model=True, gened_proto=gened.cuda(), base_class_coding=coding[:7].cuda(), novel_class_coding=coding[7:].cuda())
    with torch.no_grad():
        full, _, _ = m(x=xc, **kw)
        solo, _, _ = m(x=xc[5:6].contiguous(), **kw)
        pair, _, _ = m(x=xc[4:6].contiguous(), **kw)
    assert full.shape == (B, 13, N)
    assert torch.equal(full[5], solo[0]) and torch.equal(full[5], pair[1]), "a block's result must not depend on its batch"
    sd = golden_sd("gfs_s3dis_weights")
    sample = [0, 17, 31]
    with torch.no_grad():
        ref, f = O.forward_eval(sd, gp, x[sample], gened, coding[:7], coding[7:], 7, 1.2)
    got = full[sample].cpu()
    agree = float((got.argmax(1) == ref.argmax(1)).float().mean())
    same_gw = float((m._features(xc[sample].contiguous())[1].cpu().long() == f["assignment"]).float().mean())
    print(f"full-size label agreement on {len(sample)} x {N} points: {agree:.5f}; GW assignment agreement {same_gw:.5f}")
    assert agree >= 0.999 and same_gw >= 0.99


def test_config4_kmeans_shard_properties():
    """configs[3] per-GPU shard (4 M / 8 = 500 k points, 150 centroids): checksums and sampled fp64 verification"""
    from gfs3d import ops
    n, D, K = 500000, 192, 150
    g = torch.Generator(device="cuda").manual_seed(1)
    cent = torch.randn(K, D, device="cuda", generator=g)
    X = cent[torch.randint(0, K, (n,), device="cuda", generator=g)] + 0.35 * torch.randn(n, D, device="cuda", generator=g)
    C = X[torch.randperm(n, device="cuda", generator=g)[:K]].clone()
    ct = torch.zeros(D, 152, device="cuda")
    ct[:, :K] = C.t()
    labels = ops.kmeans_assign(X.t().contiguous(), ct, K)
    sums, counts = ops.kmeans_accumulate(X, labels, K)
    assert int(counts.sum()) == n and int(labels.min()) >= 0 and int(labels.max()) < K
    assert torch.equal(counts, torch.bincount(labels.long(), minlength=K))
    # checksum of checksums: the centroid sums add up to the sum of all points
    assert rel_err(sums.sum(0).cpu(), X.double().sum(0).cpu()) <= 1e-5
    # sampled points: the chosen centroid is the nearest one (fp64), up to fp32 rounding of the score
    s = torch.randperm(n, device="cuda", generator=g)[:4096]
    d = ((X[s].double()[:, None, :] - C.double()[None]) ** 2).sum(-1)
    best = d.min(1).values
    chosen = d[torch.arange(len(s), device="cuda"), labels[s].long()]
    assert bool((chosen <= best + 1e-4 * best.abs() + 1e-4).all())
    # bit-exact against the pinned-order oracle on a slice
    ref = O.kmeans_assign_exact(X[:20000].cpu().numpy(), C.cpu().numpy())
    assert np.array_equal(labels[:20000].cpu().numpy(), ref)


def test_both_knn_kernels_give_the_same_model_output_bit_for_bit(golden_sd):
    """End to end at B = 16 x 2048: the tensor-core filtered kNN and the all-fp32 kNN produce the same three graphs on the
    model's own (dynamic) feature spaces, hence bit-identical logits -- including the collapsed layer-3 features of an
    untrained network, where many candidates sit inside the filter's error bound."""
    from gfs3d import ops
    from gfs3d.synthetic import synthetic_blocks
    x = synthetic_blocks(16, 2048, seed=777).cuda()
    g = torch.Generator().manual_seed(12)
    gp = torch.randn(150, 192, generator=torch.Generator().manual_seed(7))
    gened = torch.nn.functional.normalize(torch.randn(13, 128, generator=g), dim=1)
    coding = (torch.rand(13, 150, generator=g) < 0.3).float()
    m = _model(golden_sd, gp)
    kw = dict(y=None, eval_model=True, gened_proto=gened.cuda(), base_class_coding=coding[:7].cuda(), novel_class_coding=coding[7:].cuda())
    outs = {}
    saved = ops.KNN_IMPL
    try:
        for impl in ("exact", "tc"):
            ops.KNN_IMPL = impl
            with torch.no_grad():
                outs[impl] = m(x=x, **kw)[0].clone()
    finally:
        ops.KNN_IMPL = saved
    assert torch.equal(outs["exact"], outs["tc"])


def _gw_flips_are_near_ties(assign_gpu, f, gp, tol):
    """every point whose GW assignment differs from the oracle's must be a near-tie of the oracle's two best words: the gap of
    10*cos between them is below `tol` (the bf16 conv2 feature noise); returns (agreement, worst gap among the flips)"""
    same = assign_gpu.long() == f["assignment"]
    if bool(same.all()):
        return 1.0, 0.0
    top2 = (10.0 * f["cos"]).topk(2, dim=1).values          # (B, 2, N)
    gap = top2[:, 0] - top2[:, 1]
    return float(same.float().mean()), float(gap[~same].max())


@pytest.mark.parametrize("N,B,k", [(4096, 2, 20), (8192, 1, 20), (4096, 1, 40)])
def test_config5_large_blocks_end_to_end_vs_oracle(golden_sd, N, B, k):
    """BASELINE.json configs[4] shapes (N = 4096 / 8192 points per block, k = 20 / 40): full eval forward against the oracle on
    the same inputs -- labels >= 99.9 %, logits within the bf16 tolerance where the GW assignment agrees, every GW flip a
    near-tie of the oracle's own two best words"""
    from model.capl import mpti_net_Point_GeoAsWeight_v2
    from gfs3d.synthetic import synthetic_blocks
    args = SimpleNamespace(edgeconv_widths=[[64, 64]] * 3, dgcnn_mlp_widths=[512, 256], pc_in_dim=9, dgcnn_k=k,
                           base_widths=[128, 64], output_dim=64, eval_weight=1.2)
    gp = torch.randn(150, 192, generator=torch.Generator().manual_seed(7))
    m = mpti_net_Point_GeoAsWeight_v2(classes=13, criterion=torch.nn.CrossEntropyLoss(ignore_index=255), args=args, base_num=7,
                                      gp=gp.cuda(), energy=0.9)
    sd = golden_sd("gfs_s3dis_weights")
    m.load_state_dict(sd)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(13)
    gened = torch.nn.functional.normalize(torch.randn(13, 128, generator=g), dim=1)
    coding = (torch.rand(13, 150, generator=g) < 0.3).float()
    x = synthetic_blocks(B, N, seed=31 + N)
    with torch.no_grad():
        got, _, _ = m(x=x.cuda(), y=None, eval_model=True, gened_proto=gened.cuda(), base_class_coding=coding[:7].cuda(),
                      novel_class_coding=coding[7:].cuda())
        assign = m._features(x.cuda())[1].cpu()
        ref, f = O.forward_eval(sd, gp, x, gened, coding[:7], coding[7:], 7, 1.2, k=k)
    agree = float((got.cpu().argmax(1) == ref.argmax(1)).float().mean())
    a_agree, worst_gap = _gw_flips_are_near_ties(assign, f, gp, 0.1)
    same = (assign.long() == f["assignment"]).unsqueeze(1).expand_as(ref)
    abs_err = float((got.cpu() - ref)[same].abs().max())
    err = abs_err / float(ref.abs().max())
    # label flips: with random-init weights the class logits of a point are nearly tied (every point sees almost the same
    # feature vector), so a flip is legitimate exactly where the oracle's own two best class logits are closer than twice the
    # logit error; every point that is decidable at that resolution must agree
    t2 = ref.topk(2, dim=1).values
    decidable = (t2[:, 0] - t2[:, 1]) > 2.0 * abs_err
    wrong = got.cpu().argmax(1) != ref.argmax(1)
    print(f"N={N} k={k}: label agreement {agree:.5f} ({float(decidable.float().mean()):.4f} of the points decidable at the logit error "
          f"{abs_err:.2e}; flips among them: {int((wrong & decidable).sum())}), GW assignment agreement {a_agree:.5f} (largest logit gap "
          f"among flips {worst_gap:.4f}), logits rel err {err:.3e}")
    assert err <= 2e-2
    assert int((wrong & decidable).sum()) == 0, "a label flip on a point whose class logits are NOT a near-tie"
    assert agree >= 0.99
    assert a_agree >= 0.995 and worst_gap <= 0.1, "a GW flip that is not a near-tie of the oracle's two best words"


def test_knn_filter_error_bound_on_all_three_layers_of_the_bench_model(golden_sd):
    """the tensor-core filter's error bound is measured, not proven -> keep it measured where it matters: the three
    kNN inputs of the benchmark model (raw blocks, EdgeConv-1 features, EdgeConv-2 features) at B = 32 x 2048.  For every pair
    the filter value must stay inside HALF of the pair bound  a_i + a_j  that the exactness argument of csrc/knn_tc.cu uses,
    no tile may need the repair pass on layers 1-2, and the indices must equal the all-fp32 kernel's bit for bit."""
    from gfs3d import ops
    from gfs3d.synthetic import synthetic_blocks
    B, N, k = 32, 2048, 20
    x = synthetic_blocks(B, N, seed=777).cuda()
    gp = torch.randn(150, 192, generator=torch.Generator().manual_seed(7))
    m = _model(golden_sd, gp)
    assert m.encoder.return_edgeconvs
    with torch.no_grad():
        ecs, _ = m.encoder(x)
    assert len(ecs) == 3
    layers = [x, ecs[0].contiguous(), ecs[1].contiguous()]
    for li, f in enumerate(layers):
        C = f.shape[1]
        idx, filt, flags = ops.knn_tc_diag(f, k)
        exact = ops.knn(f, k, impl="exact")
        worst = 0.0
        for b in range(0, B, 8):                                   # fp64 check of four blocks per layer
            fd = f[b].double()
            xx = (fd * fd).sum(0)
            xx_fl = torch.zeros(N, dtype=torch.float32, device=f.device)      # the pinned fp32 chain's |x_j|^2 (fma = fp64 sum rounded)
            for c in range(C):
                xx_fl = (xx_fl.double() + fd[c] * fd[c]).float()
            d_true = 2.0 * (fd.t() @ fd) - xx[:, None] - xx_fl.double()[None, :]
            f32 = f[b]
            mu = 0.25 * ((f32[:, 0] + f32[:, N // 4]) + (f32[:, N // 2] + f32[:, 3 * (N // 4)]))
            xc = fd - mu.double()[:, None]
            cc = (xc * xc).sum(0)
            aw = (torch.arange(C, 0, -1, dtype=torch.float64, device=fd.device)[:, None] * fd * fd).sum(0)
            a = 2.0 ** -15 * cc + 2.0 ** -25 * aw + 6 * 2.0 ** -25 * xx
            resid = 2.0 * (filt[b, :, :N].double() - a[None, :]) - d_true
            err = (resid - resid.median(dim=1, keepdim=True).values).abs()
            worst = max(worst, float((err / (2.0 * (a[:, None] + a[None, :]))).max()))
        print(f"layer {li} (C={C}): max filter error / pair bound = {worst:.4f}; tiles sent to the repair pass: {int(flags.sum())} of {flags.numel()}")
        assert worst < 0.5
        if li < 2:
            assert int(flags.sum()) == 0
        assert torch.equal(idx, exact)

"""GPU suite: the tensor-core filtered kNN (gfs_knn_tc_f32) must return the bits of the all-fp32 kernel and of the
pinned-order oracle, and its filter's error must stay inside the margin the proof of exactness assumes."""
import pytest
import torch

from oracle import gfs_oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    from gfs3d import ops
    return ops


def _sqnorm_fp32_chain(x32):
    """|x_j|^2 as the pinned fp32 fma chain computes it (x32: (C, N) fp32): s = fl32(s + v * v) per channel.  v * v is exact in
    fp64, so rounding the fp64 sum to fp32 reproduces the fused multiply-add."""
    s = torch.zeros(x32.shape[1], dtype=torch.float32, device=x32.device)
    for c in range(x32.shape[0]):
        v = x32[c].double()
        s = (s.double() + v * v).float()
    return s.double()


def _both(x, k):
    ops = _ops()
    a = ops.knn(x, k, return_dist=True, impl="exact")
    b = ops.knn(x, k, return_dist=True, impl="tc")
    torch.cuda.synchronize()
    return a, b


@pytest.mark.parametrize("B,C,N,k", [(4, 64, 2048, 20), (4, 9, 2048, 20), (2, 64, 1000, 20), (3, 6, 300, 16),
                                     (2, 48, 516, 20), (1, 17, 260, 1), (2, 64, 4096, 20), (1, 12, 8192, 20)])
def test_tc_equals_exact_kernel_bit_for_bit(B, C, N, k):
    g = torch.Generator().manual_seed(B * 1000 + C * 10 + N)
    x = torch.randn(B, C, N, generator=g).cuda() * 0.7
    (ia, da), (ib, db) = _both(x, k)
    assert torch.equal(ia, ib), "indices differ between the tensor-core filtered and the all-fp32 kernel"
    assert torch.equal(da, db), "distances differ"


def test_tc_equals_exact_kernel_on_random_shapes():
    """30 seeded random (B, C, N, k) with clustered / duplicated / anisotropic data: the two kernels must agree bit for bit"""
    rng = torch.Generator().manual_seed(2024)
    for trial in range(30):
        B = int(torch.randint(1, 5, (1,), generator=rng))
        C = int(torch.randint(1, 65, (1,), generator=rng))
        N = 4 * int(torch.randint(6, 600, (1,), generator=rng))
        k = int(torch.randint(1, min(20, N) + 1, (1,), generator=rng))
        x = torch.randn(B, C, N, generator=rng)
        kind = trial % 4
        if kind == 1:                                   # clusters: many near neighbours per point
            cent = torch.randn(B, C, 8, generator=rng) * 5
            x = cent[:, :, torch.randint(0, 8, (N,), generator=rng)] + 0.05 * x
        elif kind == 2:                                 # duplicates (sampling with replacement)
            src = torch.randint(0, N, (N // 3,), generator=rng)
            dst = torch.randint(0, N, (N // 3,), generator=rng)
            x[:, :, dst] = x[:, :, src]
        elif kind == 3:                                 # far from the origin, anisotropic
            x = x * torch.logspace(-2, 1, C).view(1, C, 1) + 30.0
        (ia, da), (ib, db) = _both(x.cuda().contiguous(), k)
        assert torch.equal(ia, ib) and torch.equal(da, db), f"trial {trial}: B={B} C={C} N={N} k={k} kind={kind}"


def test_tc_on_edgeconv_like_features_full_batch():
    """post-LeakyReLU, correlated 64-channel features (what the second and third kNN of the backbone see)"""
    g = torch.Generator().manual_seed(5)
    B, N = 8, 2048
    pos = torch.rand(B, 3, N, generator=g)
    w = torch.randn(64, 3, generator=g)
    f = torch.nn.functional.leaky_relu(torch.einsum("oc,bcn->bon", w, pos) + 0.1 * torch.randn(B, 64, N, generator=g), 0.2)
    (ia, da), (ib, db) = _both(f.cuda().contiguous(), 20)
    assert torch.equal(ia, ib) and torch.equal(da, db)
    ref = O.knn_exact(f[:1], 20)
    assert torch.equal(ib[:1].cpu(), ref)


@pytest.mark.parametrize("case", ["zeros", "offset", "huge", "tiny", "one_hot"])
def test_tc_adversarial_inputs(case):
    g = torch.Generator().manual_seed(11)
    B, C, N, k = 2, 9, 512, 20
    x = torch.rand(B, C, N, generator=g)
    if case == "zeros":
        x.zero_()                                   # every distance ties: all rows go to the exact repair pass
    elif case == "offset":
        x += 1000.0                                 # |x|^2 >> neighbour spacing: the margin swallows many candidates
    elif case == "huge":
        x *= 1e15
    elif case == "tiny":
        x *= 1e-18
    elif case == "one_hot":
        x.zero_()
        x[:, 0, ::2] = 1.0                          # two clusters of identical points
    idx_ref, d_ref = O.knn_exact(x, k, return_dist=True)
    (ia, da), (ib, db) = _both(x.cuda(), k)
    assert torch.equal(ib.cpu(), idx_ref), "tc differs from the oracle"
    assert torch.equal(ia, ib) and torch.equal(da, db)
    assert torch.equal(db.cpu(), d_ref)


@pytest.mark.parametrize("C,N,scale,offset", [(9, 2048, 1.0, 0.0), (64, 2048, 1.0, 0.0), (64, 1024, 37.0, 0.0),
                                              (33, 512, 1e-3, 0.0), (64, 2048, 1.0, 5.0), (9, 1024, 0.05, 3.0)])
def test_filter_error_is_inside_the_margin(C, N, scale, offset):
    """The proof of exactness in knn_tc.cu needs |v(i,j) - D(i,j)| <= a_i + a_j for the tensor-core value v, with
    a = 2^-15 |x~|^2 + 2^-24 sum_c (C-c) x_c^2 + 6 2^-25 |x|^2 per point.  The diagnostic entry dumps u = v + a_j; in units of d = 2 D + const_i the
    requirement reads |2 (u - a_j) - d - const_i| <= 2 (a_i + a_j).  Measure it: it must stay below HALF of that."""
    ops = _ops()
    g = torch.Generator().manual_seed(C + N)
    x = (torch.randn(1, C, N, generator=g) * scale + offset).cuda()
    idx, filt, flags = ops.knn_tc_diag(x, 20)
    torch.cuda.synchronize()
    xd = x[0].double()
    xx = (xd * xd).sum(0)
    # real-arithmetic d(i, j), except that candidate j's square norm is the COMPUTED one (the filter constant carries its
    # known rounding, see knn_prep_kernel); the row's own term is removed with the median below
    d_true = 2.0 * (xd.t() @ xd) - xx[:, None] - _sqnorm_fp32_chain(x[0])[None, :]
    x32 = x[0]
    mu = 0.25 * ((x32[:, 0] + x32[:, N // 4]) + (x32[:, N // 2] + x32[:, 3 * (N // 4)]))       # the kernel's shift
    xc = xd - mu.double()[:, None]
    cc = (xc * xc).sum(0)
    aw = (torch.arange(C, 0, -1, dtype=torch.float64, device=xd.device)[:, None] * xd * xd).sum(0)
    a = 2.0 ** -15 * cc + 2.0 ** -25 * aw + 6 * 2.0 ** -25 * xx
    resid = 2.0 * (filt[0, :, :N].double() - a[None, :]) - d_true        # = |x~_i|^2 + error
    err = (resid - resid.median(dim=1, keepdim=True).values).abs()
    ratio = float((err / (2.0 * (a[:, None] + a[None, :]))).max())
    print(f"C={C} N={N} scale={scale} offset={offset}: max filter error / pair bound = {ratio:.4f}, "
          f"repaired tiles {int(flags.sum())}")
    assert ratio < 0.5, "the filter's error is too close to the bound the exactness proof assumes"
    assert int(flags.sum()) == 0, "random data must not need the repair pass"
    assert torch.equal(idx.cpu(), O.knn_exact(x.cpu(), 20))


def test_tie_flood_rows_are_repaired_not_guessed():
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 9, 1024, generator=g)
    x[:, :, 100:400] = x[:, :, 100:101]             # 300 identical points: their rows cannot be decided by the filter
    idx, filt, flags = ops.knn_tc_diag(x.cuda(), 20)
    assert int(flags.sum()) > 0
    assert torch.equal(idx.cpu(), O.knn_exact(x, 20))


def test_moderate_tie_floods_stay_on_the_fast_path():
    """100 identical points: their rows keep ~100 survivors (spilled to the row's list in global memory, ranked in four
    rounds by the finish kernel) - no repair pass, and still the oracle's answer (ties -> ascending index)"""
    ops = _ops()
    g = torch.Generator().manual_seed(4)
    x = torch.rand(2, 64, 1024, generator=g)
    x[:, :, 300:400] = x[:, :, 300:301]
    idx, filt, flags = ops.knn_tc_diag(x.cuda(), 20)
    assert int(flags.sum()) == 0
    assert torch.equal(idx.cpu(), O.knn_exact(x, 20))
    (ia, da), (ib, db) = _both(x.cuda(), 20)
    assert torch.equal(ia, ib) and torch.equal(da, db)


def test_tc_rejects_what_it_does_not_build():
    ops = _ops()
    with pytest.raises(RuntimeError, match="k=41"):
        ops.knn(torch.randn(1, 9, 128, device="cuda"), 41, impl="tc")


def _sets_equal(a, b):
    return torch.equal(a.sort(dim=-1).values, b.sort(dim=-1).values)


@pytest.mark.parametrize("B,C,N,k", [(4, 64, 2048, 20), (4, 9, 2048, 20), (2, 64, 1000, 20), (3, 6, 300, 16), (1, 17, 260, 1),
                                     (2, 64, 4096, 20), (2, 9, 320, 40), (1, 64, 1024, 40), (2, 9, 20, 20), (1, 3, 4, 1)])
def test_set_mode_returns_the_same_neighbour_sets(B, C, N, k):
    """gfs_knn_tc_set_f32 (bounds-classified, exact arithmetic on the undecided band only) == the ordered result as sets"""
    ops = _ops()
    g = torch.Generator().manual_seed(B * 1000 + C * 10 + N + 1)
    x = (torch.randn(B, C, N, generator=g) * 0.7).cuda()
    ref = ops.knn(x, k, impl="exact")
    got = ops.knn(x, k, impl="tc", ordered=False)
    torch.cuda.synchronize()
    assert _sets_equal(ref, got)
    assert torch.equal(O.knn_exact(x[:1].cpu(), k).sort(dim=-1).values, got[:1].cpu().sort(dim=-1).values)


def test_set_mode_on_hard_inputs():
    """duplicates, clusters, far-from-origin data, post-LeakyReLU features, tie floods: the set must still be the exact one
    (exact ties at the k-th place: the lower index wins, as in the ordered kernels)"""
    ops = _ops()
    rng = torch.Generator().manual_seed(77)
    for trial in range(16):
        B, C = 2, [9, 64, 33, 12][trial % 4]
        N = 4 * int(torch.randint(40, 600, (1,), generator=rng))
        k = [20, 20, 7, 40][trial // 4]
        x = torch.randn(B, C, N, generator=rng)
        kind = trial % 4
        if kind == 1:
            cent = torch.randn(B, C, 8, generator=rng) * 5
            x = cent[:, :, torch.randint(0, 8, (N,), generator=rng)] + 0.05 * x
        elif kind == 2:
            src = torch.randint(0, N, (N // 3,), generator=rng)
            dst = torch.randint(0, N, (N // 3,), generator=rng)
            x[:, :, dst] = x[:, :, src]
        elif kind == 3:
            x = torch.nn.functional.leaky_relu(x, 0.2) * torch.logspace(-2, 1, C).view(1, C, 1) + 30.0
        x = x.cuda().contiguous()
        ref = ops.knn(x, k, impl="exact")
        got = ops.knn(x, k, impl="tc", ordered=False)
        assert _sets_equal(ref, got), f"trial {trial}: C={C} N={N} k={k} kind={kind}"
    x = torch.rand(1, 9, 1024, generator=rng)
    x[:, :, 100:400] = x[:, :, 100:101]             # tie flood: repaired by the exact kernel (which writes ordered rows)
    assert _sets_equal(ops.knn(x.cuda(), 20, impl="exact"), ops.knn(x.cuda(), 20, impl="tc", ordered=False))


def test_tc_large_common_offset_c64():
    """all-positive 64-channel features far from the origin relative to their spread: the fp32 term of the bound
    dominates; the result must still be the exact one, through the filter or through the repair pass"""
    g = torch.Generator().manual_seed(21)
    x = (torch.rand(2, 64, 1024, generator=g) * 0.05 + 40.0)
    (ia, da), (ib, db) = _both(x.cuda(), 20)
    assert torch.equal(ia, ib) and torch.equal(da, db)
    assert torch.equal(ib[:1].cpu(), O.knn_exact(x[:1], 20))


@pytest.mark.parametrize("B,C,N,k,first", [(5, 64, 1024, 20, 2), (3, 9, 516, 16, 1), (4, 33, 2048, 40, 3), (2, 64, 260, 20, 5)])
def test_two_chains_equal_one_chain(B, C, N, k, first):
    """gfs_knn_tc_set_split: the blocks processed as two chains side by side (second chain on the library's side stream)
    give the bits of the single chain -- ordered output with distances, set output, tie-flood repair included -- and the
    call can be captured into a CUDA graph with the fork / join inside."""
    from gfs3d._lib import lib
    ops = _ops()
    g = torch.Generator().manual_seed(B * 77 + N)
    x = torch.randn(B, C, N, generator=g)
    x[B - 1, :, : N // 2] = 0.25                      # a tie flood in the last block: repair pass inside the second chain
    x = x.cuda()
    try:
        assert lib().gfs_knn_tc_set_split(0) == 0
        assert lib().gfs_knn_tc_chains(B, C, N) == 1
        i1, d1 = ops.knn(x, k, return_dist=True, impl="tc")
        s1 = ops.knn(x, k, impl="tc", ordered=False)
        assert lib().gfs_knn_tc_set_split(first) == 0
        assert lib().gfs_knn_tc_chains(B, C, N) == 2
        n0 = ops.LAUNCHES
        i2, d2 = ops.knn(x, k, return_dist=True, impl="tc")
        assert ops.LAUNCHES - n0 == 8
        s2 = ops.knn(x, k, impl="tc", ordered=False)
        torch.cuda.synchronize()
        assert torch.equal(i1, i2) and torch.equal(d1, d2)
        assert torch.equal(s1.sort(dim=2).values, s2.sort(dim=2).values)
        assert torch.equal(s2.sort(dim=2).values, i1.sort(dim=2).values)
        # graph capture: the side stream is forked from and joined back into the capturing stream
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ops.knn(x, k, impl="tc", ordered=False)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            s3 = ops.knn(x, k, impl="tc", ordered=False)
        s3.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(s3.sort(dim=2).values, s1.sort(dim=2).values)
    finally:
        lib().gfs_knn_tc_set_split(-1)

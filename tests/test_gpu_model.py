"""GPU suite: head kernels and the drop-in modules vs the oracle and the committed reference fixtures."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import gfs_oracle as O
from parity import knn_classify_mismatches, rel_err

pytestmark = pytest.mark.gpu

TOL_FP32 = 1e-3     # north_star: fp32 paths
TOL_BF16 = 2e-2     # north_star: bf16 GEMM paths


def _ops():
    from gfs3d import ops
    return ops


def _args(k=20, eval_weight=1.2):
    return SimpleNamespace(edgeconv_widths=[[64, 64]] * 3, dgcnn_mlp_widths=[512, 256], pc_in_dim=9, dgcnn_k=k,
                           base_widths=[128, 64], output_dim=64, eval_weight=eval_weight)


@pytest.mark.parametrize("M_shape,K,Nout,act", [((2, 256), 192, 512, 1), ((1, 2048), 512, 256, 1), ((3, 200), 256, 128, 2),
                                                ((2, 128), 128, 64, 0), ((1, 320), 256, 192, 0), ((2, 256), 384, 128, 1)])
def test_linear_tcgen05_vs_fp64(M_shape, K, Nout, act):
    ops = _ops()
    B, N = M_shape
    g = torch.Generator().manual_seed(K + Nout)
    x = torch.randn(B, K, N, generator=g)
    w = torch.randn(Nout, K, generator=g) / K ** 0.5
    s = 1 + 0.3 * torch.randn(Nout, generator=g)
    t = 0.2 * torch.randn(Nout, generator=g)
    xa = ops.new_act(B * N, K // 64, "cuda")
    ops.cm_to_act(x.cuda(), xa, 0)
    dense = ops.act_to_dense(xa, B * N).float().cpu()                      # the bf16 values the kernel really sees
    assert rel_err(dense.reshape(B, N, K).permute(0, 2, 1), x) <= 8e-3
    wp = ops.pack_weight(w.cuda(), s.cuda())
    y_act = ops.new_act(B * N, Nout // 64 + 1, "cuda")
    y_cm = torch.empty(B, Nout, N, device="cuda")
    ops.linear(xa, 0, K // 64, wp, t.cuda(), Nout, act, B, N, y_act=y_act, y_kb0=1, y_cm=y_cm)
    torch.cuda.synchronize()
    wb = (w * s[:, None]).bfloat16().double()
    ref = dense.double() @ wb.t() + t.double()
    ref = {0: ref, 1: F.leaky_relu(ref, 0.2), 2: F.relu(ref)}[act]
    ref = ref.reshape(B, N, Nout).permute(0, 2, 1)
    assert rel_err(y_cm.cpu(), ref) <= 1e-4, "tcgen05 GEMM differs from fp64 on identical bf16 operands"
    got = ops.act_to_dense(y_act, B * N).float().cpu()
    assert float(got[:, :64].abs().max()) == 0
    assert rel_err(got[:, 64:].reshape(B, N, Nout).permute(0, 2, 1), y_cm.cpu()) <= 8e-3
    # against the un-rounded fp32 math: this is the 2e-2 bf16 tolerance of north_star
    full = torch.einsum("ok,bkn->bon", (w * s[:, None]).double(), x.double()) + t.double().view(1, -1, 1)
    full = {0: full, 1: F.leaky_relu(full, 0.2), 2: F.relu(full)}[act]
    assert rel_err(y_cm.cpu(), full) <= TOL_BF16


@pytest.mark.parametrize("B,N,G", [(2, 256, 150), (1, 2048, 150), (2, 128, 180), (1, 64, 64)])
def test_gw_projection_vs_oracle(B, N, G):
    ops = _ops()
    g = torch.Generator().manual_seed(G + N)
    ec = torch.randn(B, 192, N, generator=g).abs() * 0.3
    gp = torch.randn(G, 192, generator=g)
    cos = torch.matmul(F.normalize(gp.double(), dim=1).unsqueeze(0), F.normalize(ec.double(), dim=1))
    ref = torch.softmax(10 * cos, dim=1)
    Gp = (G + 63) // 64 * 64
    gp_l2t = torch.zeros(192, Gp)
    gp_l2t[:, :G] = F.normalize(gp, dim=1).t()
    act = ops.new_act(B * N, Gp // 64 + 1, "cuda")
    assign, cm = ops.gw_project(ec.cuda(), gp_l2t.cuda(), G, cosine_act=act, kb0=1, want_cm=True)
    torch.cuda.synchronize()
    assert rel_err(cm.cpu(), ref) <= TOL_FP32
    a_ref = ref.argmax(1)
    agree = (assign.cpu().long() == a_ref).float().mean()
    # disagreements must be fp32 near-ties of the two best cosines
    bad = (assign.cpu().long() != a_ref).nonzero()
    for b, n in bad.tolist():
        top2 = cos[b, :, n].topk(2).values
        assert float(top2[0] - top2[1]) < 1e-5
    assert agree >= 0.999
    dense = ops.act_to_dense(act, B * N).float().cpu()
    assert rel_err(dense[:, 64:64 + G].reshape(B, N, G).permute(0, 2, 1), ref) <= 8e-3
    assert float(dense[:, 64 + G:].abs().max()) == 0 if Gp > G else True


@pytest.mark.parametrize("B,N,CLS,per_batch", [(2, 256, 13, False), (2, 256, 13, True), (1, 2048, 21, True), (3, 100, 22, False)])
def test_cos_logits_and_pool_vs_oracle(B, N, CLS, per_batch):
    ops = _ops()
    g = torch.Generator().manual_seed(CLS * N)
    feat = torch.randn(B, 128, N, generator=g)
    proto = torch.randn(*((B, CLS, 128) if per_batch else (CLS, 128)), generator=g)
    ref = O.get_pred(feat.double(), proto.double())
    got = ops.cos_logits(feat.cuda(), F.normalize(proto, dim=-1).cuda())
    assert rel_err(got.cpu(), ref) <= TOL_FP32
    G = 150
    coding = (torch.rand(CLS, G, generator=g) < 0.3).float()
    assign = torch.randint(0, G, (B, N), generator=g).int()
    onehot = F.one_hot(assign.long(), G).transpose(2, 1).double()
    refw = ref * O.get_gp_weight(coding.double(), onehot, 1.7)
    gotw = ops.cos_logits(feat.cuda(), F.normalize(proto, dim=-1).cuda(), coding.cuda(), assign.cuda(), 1.7)
    assert rel_err(gotw.cpu(), refw) <= TOL_FP32
    if not per_batch:
        pp_ref = O.post_refine_proto_v2(proto.double(), feat.double())
        pool = ops.softmax_pool(got, feat.cuda())
        pred = torch.softmax(ref, dim=2) @ feat.double().transpose(1, 2)
        assert rel_err(pool.cpu(), pred) <= TOL_FP32
        assert pp_ref.shape == (B, CLS, 128)


def _load_dgcnn(golden_sd, k=20):
    from model.dgcnn import DGCNN
    m = DGCNN([[64, 64]] * 3, [512, 256], 9, k=k, return_edgeconvs=True)
    missing = m.load_state_dict(golden_sd("dgcnn_weights"), strict=True)      # state-dict contract
    return m.cuda().eval()


@pytest.mark.parametrize("name", ["dgcnn_b2_n256", "dgcnn_dup_b1_n256", "dgcnn_b1_n2048", "dgcnn_b1_n320_k40"])
def test_dgcnn_dropin_vs_reference_fixture(golden, golden_sd, name):
    """BASELINE.json configs[0] (N=2048) and small cases: the drop-in DGCNN against outputs of the real reference."""
    g = golden(name)
    m = _load_dgcnn(golden_sd, k=int(g["k"]))
    x = torch.from_numpy(g["x"])
    s = int(g["subsample"])
    with torch.no_grad():
        ecs, out = m(x.cuda())
    ec = torch.cat(ecs, 1).cpu()
    e1 = rel_err(ec[:, :64, ::s], torch.from_numpy(g["ec"])[:, :64])
    e = rel_err(ec[:, :, ::s], torch.from_numpy(g["ec"]))
    eo = rel_err(out.cpu()[:, :, ::s], torch.from_numpy(g["out"]))
    print(f"{name}: rel err layer1 {e1:.2e}  edgeconv123 {e:.2e}  mlp out {eo:.2e}")
    assert e1 <= TOL_BF16 and e <= TOL_BF16 and eo <= TOL_BF16
    # layer-0 neighbour sets vs the reference's own topk (input is bit-identical, so only fp32 near-ties may differ)
    from gfs3d import ops
    idx = ops.knn(x.cuda(), int(g["k"])).cpu().numpy()
    n, near, real = knn_classify_mismatches(x, idx, g["idx0"].astype(np.int64), int(g["k"]))
    print(f"{name}: kNN rows differing from reference topk: {n} (near-tie {near}, real {real}) of {x.shape[0] * x.shape[2]}")
    assert real == 0


def test_dgcnn_features_with_pinned_graph(golden, golden_sd):
    """decouple kNN near-ties from feature parity: feed the oracle the neighbour sets the CUDA path used"""
    from gfs3d import ops
    g = golden("dgcnn_b2_n256")
    sd = golden_sd("dgcnn_weights")
    m = _load_dgcnn(golden_sd)
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        ecs, out = m(x.cuda())
        idx_list = [ops.knn(t.contiguous(), 20).cpu() for t in (x.cuda(), ecs[0], ecs[1])]
        o_ecs, o_out, _ = O.dgcnn_forward(sd, x, 20, "", idx_list=idx_list)
    assert rel_err(torch.cat(ecs, 1).cpu(), torch.cat(o_ecs, 1)) <= TOL_BF16
    assert rel_err(out.cpu(), o_out) <= TOL_BF16


@pytest.mark.parametrize("name,wname,classes,base_num,G", [("gfs_s3dis_b2_n256", "gfs_s3dis_weights", 13, 7, 150),
                                                           ("gfs_scannet_b2_n128", "gfs_scannet_weights", 21, 15, 180)])
def test_gfs_model_eval_vs_reference_fixture(golden, golden_sd, name, wname, classes, base_num, G):
    from model.capl import mpti_net_Point_GeoAsWeight_v2
    g = golden(name)
    t = lambda k: torch.from_numpy(g[k])
    args = _args(eval_weight=float(g["eval_weight"]))
    m = mpti_net_Point_GeoAsWeight_v2(classes=classes, criterion=torch.nn.CrossEntropyLoss(ignore_index=255), args=args,
                                      base_num=base_num, gp=t("gp").cuda(), energy=0.9)
    m.load_state_dict(golden_sd(wname), strict=True)
    m = m.cuda().eval()
    x, y = t("x").cuda(), t("y").long().cuda()
    with torch.no_grad():
        pf, sem, oh = m.getFeatures(x)
        logits, gp_acc, gp_nacc = m(x=x, y=y, eval_model=True, gened_proto=t("gened_proto").cuda().unsqueeze(0).repeat(8, 1, 1),
                                    base_class_coding=t("base_class_coding").cuda(), novel_class_coding=t("novel_class_coding").cuda())
        fg_feat, fg_gp = m.Get_Fg_Feat(x[:1], (y[:1] == 1).long())
    e_pf, e_sem, e_lg = rel_err(pf.cpu(), t("point_feat")), rel_err(sem.cpu(), t("semantic_feat")), rel_err(logits.cpu(), t("logits"))
    a_agree = float((oh.argmax(1).cpu().numpy() == g["assignment"]).mean())
    l_agree = float((logits.argmax(1).cpu().numpy() == g["logits"].argmax(1)).mean())
    print(f"{name}: rel err point_feat {e_pf:.2e} semantic {e_sem:.2e} logits {e_lg:.2e}; "
          f"GW assignment agreement {a_agree:.5f}; label agreement {l_agree:.5f}")
    assert e_pf <= TOL_BF16 and e_sem <= TOL_BF16 and e_lg <= TOL_BF16
    assert l_agree >= 0.999, "predicted per-point labels must agree with the reference on >= 99.9 % of points"
    assert a_agree >= 0.99
    assert abs(float(gp_acc) - float(g["gp_acc"])) < 2e-2 and abs(float(gp_nacc) - float(g["gp_novel_acc"])) < 2e-2
    assert fg_feat.shape == tuple(g["fg_feat"].shape)
    assert rel_err(fg_gp.sum(0).cpu(), t("fg_gp_sum")) <= 0.05


def test_training_mode_dgcnn_runs_the_training_kernels(golden_sd):
    """model.train(): batch-statistics forward through gfs3d/train_ops.py (no PyTorch fallback), running stats are updated"""
    from gfs3d import ops
    m = _load_dgcnn(golden_sd).train()
    before = m.edge_convs[0].layer[1].running_mean.clone()
    n0 = ops.LAUNCHES
    x = torch.randn(2, 9, 128, device="cuda")
    ecs, out = m(x)
    assert ops.LAUNCHES > n0, "the training forward must run in the hand-written kernels"
    assert [tuple(e.shape) for e in ecs] == [(2, 64, 128)] * 3 and tuple(out.shape) == (2, 256, 128)
    assert out.requires_grad and torch.isfinite(out).all()
    out.sum().backward()
    assert m.conv.layer[0].weight.grad is not None and torch.isfinite(m.conv.layer[0].weight.grad).all()
    assert not torch.equal(before, m.edge_convs[0].layer[1].running_mean)
    assert int(m.edge_convs[0].layer[1].num_batches_tracked) == 1
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 9, 128))


@pytest.mark.parametrize("B,N", [(2, 256), (1, 2048), (3, 128), (2, 1024)])
def test_attention_tcgen05_vs_reference(B, N, golden_sd):
    """model/attention.py:32-48 (eval): fused q/k/v linear + flash-style tcgen05 attention vs the oracle"""
    from model.attention import SelfAttention
    ops = _ops()
    sd = golden_sd("gfs_s3dis_weights")
    att = SelfAttention(256, 64)
    att.load_state_dict({k[len("att_learner."):]: v for k, v in sd.items() if k.startswith("att_learner.")})
    att = att.cuda().eval()
    g = torch.Generator().manual_seed(N)
    sd64 = {k: v.double() for k, v in sd.items()}
    for gain in (1.0, 6.0):            # activation-scale inputs (what the model feeds it) and a peaked-softmax stress
        x = torch.randn(B, 256, N, generator=g) * gain
        with torch.no_grad():
            y = att(x.cuda())
        if gain == 1.0:
            ref = O.self_attention(sd64, "att_learner", x.double())
            tol = TOL_BF16
        else:
            # peaked softmax amplifies the bf16 rounding of q/k themselves; validate the kernel's arithmetic against
            # the same formula evaluated in fp64 on the bf16-rounded x, q, k, v it actually consumes
            xb = x.bfloat16().double()
            w = torch.cat([sd64[f"att_learner.{n}_map.weight"] for n in "qkv"]).bfloat16().double()
            qkv = torch.nn.functional.conv1d(xb, w).bfloat16().double()
            q, kk, v = qkv[:, :64], qkv[:, 64:128], qkv[:, 128:]
            attn = torch.softmax(torch.matmul(q.transpose(1, 2) / 8.0, kk), dim=-1)
            ref = torch.matmul(attn, v.transpose(1, 2)).transpose(1, 2)
            tol = 1e-2
        err = rel_err(y.cpu(), ref)
        print(f"attention B={B} N={N} gain={gain}: rel err {err:.3e}")
        assert err <= tol


def test_attention_rejects_ragged_n():
    from model.attention import SelfAttention
    att = SelfAttention(256, 64).cuda().eval()
    with pytest.raises(RuntimeError, match="multiple of 128"):
        att(torch.randn(1, 256, 320, device="cuda"))


def test_refine_proto_kernel_matches_the_reference_formula():
    """gfs_refine_proto == post_refine_proto_v2's mix + the base/novel update of model/capl.py:117-120 + F.normalize"""
    import torch.nn.functional as F
    from gfs3d import ops
    g = torch.Generator().manual_seed(17)
    B, CLS, D, base = 5, 13, 128, 7
    pp = torch.randn(B, CLS, D, generator=g)
    pp[0, 3] = -pp[0, 3].abs()                      # a row with negative cosine: w clamps to 0
    q = torch.randn(CLS, D, generator=g)
    q[3] = q[3].abs()
    gen = F.normalize(torch.randn(CLS, D, generator=g), dim=1)
    w = (F.normalize(pp, 2, -1) * F.normalize(q, 2, -1).unsqueeze(0)).sum(-1, keepdim=True)
    w = w * (w > 0).float()
    r = w * pp + (1 - w) * q.unsqueeze(0)
    r[:, :base] = r[:, :base] + gen[:base].unsqueeze(0)
    r[:, base:] = r[:, base:] * 0 + gen[base:].unsqueeze(0)
    ref = F.normalize(r, p=2, dim=-1)
    got = ops.refine_proto(pp.cuda(), q.cuda(), gen.cuda(), base).cpu()
    assert float((got - ref).abs().max()) <= 2e-6
    assert float(w[0, 3]) == 0.0


@pytest.mark.parametrize("B,N,G,kind", [(2, 256, 150, "random"), (4, 2048, 150, "random"), (2, 128, 180, "random"), (1, 128, 64, "random"),
                                        (2, 256, 150, "duplicate_words"), (1, 128, 150, "zeros")])
def test_gw_projection_tensor_core_equals_fp32_kernel(B, N, G, kind):
    """gfs_gw_project_tc: assignments identical to the all-fp32 kernel (near-ties are re-evaluated with the pinned chain),
    softmax features within the bf16 tolerance of the fp32 ones"""
    ops = _ops()
    g = torch.Generator().manual_seed(G * 7 + N)
    ec = torch.randn(B, 192, N, generator=g).abs() * 0.3
    gp = torch.randn(G, 192, generator=g)
    if kind == "duplicate_words":
        gp[G // 2:] = gp[:G - G // 2]                    # exact ties: the lower index must win
    if kind == "zeros":
        ec.zero_()
    Gp = (G + 63) // 64 * 64
    gp_l2t = torch.zeros(192, Gp)
    gp_l2t[:, :G] = F.normalize(gp, dim=1).t()
    out = {}
    for impl in ("fp32", "tc"):
        act = ops.new_act(B * N, Gp // 64 + 1, "cuda")
        assign, cm = ops.gw_project(ec.cuda(), gp_l2t.cuda(), G, cosine_act=act, kb0=1, want_cm=True, impl=impl)
        torch.cuda.synchronize()
        out[impl] = (assign.cpu(), cm.cpu(), ops.act_to_dense(act, B * N).float().cpu())
    assert torch.equal(out["fp32"][0], out["tc"][0]), f"{int((out['fp32'][0] != out['tc'][0]).sum())} assignments differ"
    if kind != "zeros":
        assert rel_err(out["tc"][1], out["fp32"][1].double()) <= 1e-3
        assert rel_err(out["tc"][2], out["fp32"][2].double()) <= 8e-3


def test_programmatic_dependent_launch_does_not_change_results(golden, golden_sd):
    """gfs_set_pdl (include/gfs3d.h): the inference kernels launched with the programmatic-serialisation attribute (each
    blocks in griddepcontrol.wait until its predecessor has completed) give bit-identical outputs to ordinary launches,
    eagerly and replayed from a CUDA graph (where the dependencies become programmatic edges)."""
    from gfs3d._lib import lib
    from model.capl import mpti_net_Point_GeoAsWeight_v2
    g = golden("gfs_s3dis_b2_n256")
    t = lambda k: torch.from_numpy(g[k])
    m = mpti_net_Point_GeoAsWeight_v2(classes=13, criterion=torch.nn.CrossEntropyLoss(ignore_index=255),
                                      args=_args(eval_weight=float(g["eval_weight"])), base_num=7, gp=t("gp").cuda(), energy=0.9)
    m.load_state_dict(golden_sd("gfs_s3dis_weights"), strict=True)
    m = m.cuda().eval()
    x = t("x").cuda().repeat(8, 1, 1)       # 16 blocks: enough CTAs for kernels to overlap their predecessors' tails
    kw = dict(y=None, eval_model=True, gened_proto=t("gened_proto").cuda().unsqueeze(0).repeat(8, 1, 1),
              base_class_coding=t("base_class_coding").cuda(), novel_class_coding=t("novel_class_coding").cuda())

    def run():
        with torch.no_grad():
            return m(x=x, **kw)[0]

    try:
        assert lib().gfs_set_pdl(0) == 0
        plain = run().clone()
        for mode in (1, 2, 3):
            assert lib().gfs_set_pdl(mode) == 0
            for _ in range(3):
                assert torch.equal(run(), plain), f"eager results differ with gfs_set_pdl({mode})"
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = run()
        for _ in range(3):
            out.zero_()
            graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(out, plain), "graph replay with programmatic edges differs from ordinary launches"
    finally:
        lib().gfs_set_pdl(3)

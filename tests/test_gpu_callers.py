"""GPU suite: the callers either side of the hot path that SURVEY.md section 8(f) ranks next -- the mIoU metric
(runs/eval.py) and the geometric-word class codings (train.py:136-241) -- on the joint-histogram kernel."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import gfs_oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    from gfs3d import ops
    return ops


@pytest.mark.parametrize("n,na,nb", [(1, 3, 5), (4099, 13, 13), (1 << 20, 22, 180), (777, 1, 1), (0, 4, 4)])
def test_joint_histogram_is_exact(n, na, nb):
    ops = _ops()
    g = torch.Generator().manual_seed(n + na)
    a = torch.randint(-2, na + 3, (n,), generator=g)          # includes labels outside the range: skipped
    b = torch.randint(-1, nb + 2, (n,), generator=g)
    if n > 10:
        a[5] = 255                                            # the loaders' "ignore" label
    ok = (a >= 0) & (a < na) & (b >= 0) & (b < nb)
    ref = torch.zeros(na, nb, dtype=torch.int64)
    ref.view(-1).index_add_(0, (a[ok] * nb + b[ok]), torch.ones(int(ok.sum()), dtype=torch.int64))
    got = ops.joint_histogram(a.cuda(), b.cuda(), na, nb)
    assert torch.equal(got.cpu(), ref)
    again = ops.joint_histogram(a.cuda(), b.cuda(), na, nb, out=got)       # accumulates
    assert torch.equal(again.cpu(), 2 * ref)


def test_joint_histogram_full_size_checksum():
    """64 M points (what a whole evaluation produces): every point lands in exactly one bin"""
    ops = _ops()
    n = 64 << 20
    a = torch.randint(0, 21, (n,), device="cuda", dtype=torch.int32)
    b = torch.randint(0, 21, (n,), device="cuda", dtype=torch.int32)
    h = ops.joint_histogram(a, b, 21, 21)
    assert int(h.sum()) == n
    assert torch.equal(h.sum(1).cpu(), torch.bincount(a.long(), minlength=21).cpu())
    assert int(torch.diagonal(h).sum()) == int((a == b).sum())


def test_joint_histogram_rejects_too_many_bins():
    ops = _ops()
    z = torch.zeros(8, dtype=torch.int32, device="cuda")
    with pytest.raises(RuntimeError, match="bins exceed"):
        ops.joint_histogram(z, z, 200, 200)
    with pytest.raises(RuntimeError):
        ops.joint_histogram(z.cpu(), z.cpu(), 4, 4)           # no CPU fallback


@pytest.mark.parametrize("name", ["metric_s3dis", "metric_scannet"])
@pytest.mark.parametrize("kind", ["numpy", "cuda"])
def test_evaluate_metric_gfs_equals_the_reference(golden, name, kind):
    """runs.eval.evaluate_metric_GFS (same import path and signature as the reference) returns the reference's floats"""
    from runs.eval import evaluate_metric_GFS
    g = golden(name)
    ncls = len(g["order"])
    pred, gt = list(g["pred"]), list(g["gt"])
    if kind == "cuda":
        pred = [torch.from_numpy(p.astype(np.int64)).cuda() for p in pred]
        gt = [torch.from_numpy(t.astype(np.int64)).cuda() for t in gt]
    lines = []
    log = SimpleNamespace(cprint=lines.append)
    mean_iou, base_iou, novel_iou, hm, ious = evaluate_metric_GFS(log, pred, gt, list(range(ncls)), g["novel"].tolist(),
                                                                  g["order"].tolist(), scannet=bool(g["scannet"]))
    assert (mean_iou, base_iou, novel_iou, hm) == (float(g["mean_iou"]), float(g["base_iou"]), float(g["novel_iou"]), float(g["hm"]))
    assert np.array_equal(ious, g["ious"])
    assert lines[0].startswith("*****Test Classes") and sum("IoU" in s for s in lines) == ncls


def test_evaluate_metric_gfs_error_behaviour_of_the_reference(golden):
    from runs.eval import evaluate_metric_GFS
    g = golden("metric_s3dis")
    log = SimpleNamespace(cprint=lambda *_: None)
    ncls = len(g["order"])
    bad = g["pred"].copy()
    bad[0, 0, 0] = ncls                                       # all_learning_order[13] -> IndexError in the reference
    with pytest.raises(IndexError):
        evaluate_metric_GFS(log, list(bad), list(g["gt"]), list(range(ncls)), g["novel"].tolist(), g["order"].tolist())
    only0 = [np.zeros((1, 8), dtype=np.int64)]               # classes that never occur: 0 / 0.0 in the reference
    with pytest.raises(ZeroDivisionError):
        evaluate_metric_GFS(log, only0, only0, list(range(ncls)), g["novel"].tolist(), g["order"].tolist())
    with pytest.raises(AssertionError):
        evaluate_metric_GFS(log, list(g["pred"]), list(g["gt"])[:-1], list(range(ncls)), g["novel"].tolist(), g["order"].tolist())


class _StubModel:
    """stands in for the GW model: returns the committed assignments (the model itself is pinned by the other suites)"""

    def __init__(self, assign, G, batch):
        self.assign, self.batch, self.i = assign, batch, 0
        self.gp = torch.zeros(G, 192)

    def eval(self):
        return self

    def _features(self, x):
        b = x.shape[0]
        a = torch.from_numpy(np.stack(self.assign[self.i:self.i + b]).astype(np.int32)).cuda()
        self.i += b
        return None, a, None


@pytest.mark.parametrize("batch", [1, 4, 12])
def test_base_class_codings_equal_the_reference(golden, batch):
    from gfs3d.coding import collect_base_class_gp_coding_sum
    g = golden("coding_s3dis")
    nb, G = int(g["num_base"]), int(g["G"])
    assign, labels = list(g["assign"]), list(g["labels"])
    loader = [(torch.zeros(len(labels[i:i + batch]), 9, labels[0].shape[0]),
               torch.from_numpy(np.stack(labels[i:i + batch]).astype(np.int64)), None) for i in range(0, len(labels), batch)]
    coding, bg = collect_base_class_gp_coding_sum(_StubModel(assign, G, batch), loader, list(range(nb)), float(g["energy"]))
    assert coding.is_cuda and coding.shape == (nb, G) and bg.shape == (G,)
    assert O.codings_equal_modulo_ties(g["freq"], coding.cpu().numpy(), g["coding"])
    oc, ob, _ = O.class_gw_codings(assign, labels, list(range(nb)), G, float(g["energy"]))
    assert np.array_equal(coding.cpu().numpy(), oc), "differs from the oracle (same tie rule: lowest index first)"
    assert float(np.abs(bg.cpu().numpy() - g["bg_coding"]).max()) <= 1e-7


def test_novel_class_codings_equal_the_reference(golden):
    from gfs3d.coding import collect_new_clsss_gp_coding_sum
    g = golden("coding_s3dis")
    G = int(g["G"])
    oh = lambda a: torch.nn.functional.one_hot(torch.from_numpy(a.astype(np.int64)), G).float().cuda()
    feats = {9: [oh(g["novel9a"]), oh(g["novel9b"])], 7: [oh(g["novel7"])]}
    freq = np.stack([torch.cat(feats[c], 0).sum(0).cpu().numpy() for c in (7, 9)])
    freq = freq / freq.sum(1, keepdims=True)
    coding = collect_new_clsss_gp_coding_sum(feats, float(g["energy"]))
    assert coding.shape == (2, G)
    assert O.codings_equal_modulo_ties(freq.astype(np.float32), coding.cpu().numpy(), g["novel_coding"])


@pytest.mark.parametrize("batch", [1, 4])
def test_support_prototypes_equal_the_reference(golden, golden_sd, batch):
    """gfs3d.coding.get_new_proto_Geo2SemProto (batched) vs train.py:240-305 run on the real reference model"""
    from gfs3d.coding import get_new_proto_Geo2SemProto
    from model.capl import mpti_net_Point_GeoAsWeight_v2
    g = golden("support_proto_s3dis")
    args = SimpleNamespace(edgeconv_widths=[[64, 64]] * 3, dgcnn_mlp_widths=[512, 256], pc_in_dim=9, dgcnn_k=20,
                           base_widths=[128, 64], output_dim=64, eval_weight=1.2)
    gp = torch.randn(150, 192, generator=torch.Generator().manual_seed(7))
    m = mpti_net_Point_GeoAsWeight_v2(classes=13, criterion=torch.nn.CrossEntropyLoss(ignore_index=255), args=args, base_num=7,
                                      gp=gp.cuda(), energy=float(g["energy"]))
    m.load_state_dict(golden_sd("gfs_s3dis_weights"))
    m = m.cuda().eval()
    x, mask, cls_id = torch.from_numpy(g["x"]), torch.from_numpy(g["mask"].astype(np.int64)), torch.from_numpy(g["cls_id"].astype(np.int64))
    loader = [(x[i:i + batch], mask[i:i + batch], cls_id[i:i + batch]) for i in range(0, x.shape[0], batch)]
    novel = sorted(set(cls_id.tolist()))
    gened, coding = get_new_proto_Geo2SemProto(loader, m, base_num=int(g["base_num"]), novel_num=len(novel), novel_class_list=novel,
                                               energy=float(g["energy"]))
    assert gened.shape == (13, 128) and coding.shape == (len(novel), 150)
    ref = torch.from_numpy(g["gened"])
    assert float((gened.cpu() - ref).abs().max()) <= 2e-2 * float(ref.abs().max()) + 1e-3      # bf16 feature path, unit-norm rows
    assert torch.allclose(gened.cpu()[:7], ref[:7], atol=1e-6)                                  # base prototypes: copied + normalised
    assert O.codings_equal_modulo_ties(g["freq"], coding.cpu().numpy(), g["coding"])

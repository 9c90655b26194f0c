"""oracle/gfs_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (functional PyTorch / numpy, state-dict driven) of the reference's hot
path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product (gfs-3dseg_gws_b200/) never does.

Parity status: the reference ships no tests or golden vectors, so this
restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build
container by tests/golden/make_golden.py (which imports /root/reference unmodified) and
committed under tests/golden/.  tests/test_oracle_golden.py re-checks that pin on CPU.

Every function cites the reference lines it restates (paths relative to the reference
repository root).  Functions take a plain ``dict[str, Tensor]`` using the reference's
state-dict key names, so they double as a check of the state-dict contract.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
_HERE = os.path.dirname(os.path.abspath(__file__))
_CLIB = None


# ----------------------------------------------------------------------------------------
# C part (pinned evaluation order; see gfs_oracle.c)
# ----------------------------------------------------------------------------------------
def build_c(force: bool = False) -> str:
    """gcc the C restatement into oracle/_build/libgfs_oracle.so (git-ignored)."""
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libgfs_oracle.so")
    src = os.path.join(_HERE, "gfs_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-mfma", "-ffp-contract=off",
               "-o", so, src, "-lm"]
        subprocess.check_call(cmd)
    return so


def _clib():
    global _CLIB
    if _CLIB is None:
        lib = ctypes.CDLL(build_c())
        lib.gfs_oracle_knn.restype = ctypes.c_int
        lib.gfs_oracle_kmeans_assign.restype = ctypes.c_int
        lib.gfs_oracle_kmeans_accumulate.restype = ctypes.c_int
        _CLIB = lib
    return _CLIB


def knn_exact(x: torch.Tensor, k: int, return_dist: bool = False):
    """model/dgcnn.py:17-23 with the pinned fma order; ties -> ascending index.

    x: (B, C, N) fp32 CPU.  Returns idx (B, N, k) int32 sorted nearest-first."""
    x = x.detach().to(torch.float32).contiguous().cpu()
    B, C, N = x.shape
    idx = np.empty((B, N, k), dtype=np.int32)
    dist = np.empty((B, N, k), dtype=np.float32)
    rc = _clib().gfs_oracle_knn(
        ctypes.c_void_p(x.data_ptr()), B, C, N, k,
        idx.ctypes.data_as(ctypes.c_void_p), dist.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise ValueError("gfs_oracle_knn: bad arguments")
    if return_dist:
        return torch.from_numpy(idx), torch.from_numpy(dist)
    return torch.from_numpy(idx)


def kmeans_assign_exact(X: np.ndarray, centers: np.ndarray, return_score: bool = False):
    """E-step of sklearn's Lloyd iteration (_k_means_lloyd.pyx:196-218) in the pinned order."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    centers = np.ascontiguousarray(centers, dtype=np.float32)
    n, D = X.shape
    K = centers.shape[0]
    labels = np.empty(n, dtype=np.int32)
    best = np.empty(n, dtype=np.float32)
    rc = _clib().gfs_oracle_kmeans_assign(
        X.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(n), D,
        centers.ctypes.data_as(ctypes.c_void_p), K,
        labels.ctypes.data_as(ctypes.c_void_p), best.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return (labels, best) if return_score else labels


def kmeans_accumulate_exact(X: np.ndarray, labels: np.ndarray, K: int):
    """M-step sums/counts (fp64 sums, ascending point order)."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    n, D = X.shape
    sums = np.zeros((K, D), dtype=np.float64)
    counts = np.zeros(K, dtype=np.int64)
    rc = _clib().gfs_oracle_kmeans_accumulate(
        X.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(n), D,
        labels.ctypes.data_as(ctypes.c_void_p), K,
        sums.ctypes.data_as(ctypes.c_void_p), counts.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return sums, counts


# ----------------------------------------------------------------------------------------
# DGCNN backbone  (model/dgcnn.py)
# ----------------------------------------------------------------------------------------
def knn_formula(x: torch.Tensor, k: int) -> torch.Tensor:
    """model/dgcnn.py:17-23 -- the reference formula with library matmul/topk (tie order and
    summation order are the library's; use knn_exact for the pinned version)."""
    gram = torch.bmm(x.transpose(1, 2), x)                 # (B,N,N)
    sq = x.pow(2).sum(dim=1, keepdim=True)                 # (B,1,N)
    neg_d = -sq - (-2.0 * gram) - sq.transpose(1, 2)
    return neg_d.topk(k, dim=-1).indices


def edge_feature(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """model/dgcnn.py:26-42 -- (B,C,N),(B,N,k) -> (B,2C,N,k) = cat(x_j - x_i, x_i)."""
    B, C, N = x.shape
    k = idx.shape[-1]
    flat = idx.reshape(B, 1, N * k).expand(B, C, N * k).long()
    nbr = x.gather(2, flat).reshape(B, C, N, k)
    ctr = x.unsqueeze(-1).expand(B, C, N, k)
    return torch.cat([nbr - ctr, ctr], dim=1)


def _bn(sd: SD, prefix: str, x: torch.Tensor, training: bool = False) -> torch.Tensor:
    """nn.BatchNorm{1,2}d, eps 1e-5, momentum 0.1 (model/dgcnn.py:54-55,73-74)."""
    return F.batch_norm(x, sd[prefix + ".running_mean"].clone(), sd[prefix + ".running_var"].clone(),
                        sd[prefix + ".weight"], sd[prefix + ".bias"], training, 0.1, 1e-5)


def conv_stack(sd: SD, prefix: str, x: torch.Tensor, n_layers: int, training: bool = False) -> torch.Tensor:
    """model/dgcnn.py:45-80 -- (Conv{1,2}d 1x1 no-bias, BN, LeakyReLU(0.2)) x n_layers."""
    conv = F.conv2d if x.dim() == 4 else F.conv1d
    for i in range(n_layers):
        x = conv(x, sd[f"{prefix}.layer.{3 * i}.weight"])
        x = _bn(sd, f"{prefix}.layer.{3 * i + 1}", x, training)
        x = F.leaky_relu(x, 0.2)
    return x


def _count_layers(sd: SD, prefix: str) -> int:
    n = 0
    while f"{prefix}.layer.{3 * n}.weight" in sd:
        n += 1
    return n


def dgcnn_forward(sd: SD, x: torch.Tensor, k: int = 20, prefix: str = "",
                  idx_list: Optional[Sequence[torch.Tensor]] = None, knn: str = "formula",
                  training: bool = False) -> Tuple[List[torch.Tensor], torch.Tensor, List[torch.Tensor]]:
    """model/dgcnn.py:113-127.  Returns (edgeconv_outputs, mlp_out, idx_used).

    idx_list pins the neighbour sets (used to decouple feature parity from kNN near-ties)."""
    n_ec = 0
    while f"{prefix}edge_convs.{n_ec}.layer.0.weight" in sd:
        n_ec += 1
    outs, used = [], []
    for i in range(n_ec):
        if idx_list is not None:
            idx = idx_list[i].long()
        elif knn == "exact":
            idx = knn_exact(x, k).long()
        else:
            idx = knn_formula(x, k)
        used.append(idx)
        p = f"{prefix}edge_convs.{i}"
        e = conv_stack(sd, p, edge_feature(x, idx), _count_layers(sd, p), training)
        x = e.max(dim=-1).values
        outs.append(x)
    p = f"{prefix}conv"
    out = conv_stack(sd, p, torch.cat(outs, dim=1), _count_layers(sd, p), training)
    return outs, out, used


# ----------------------------------------------------------------------------------------
# adjacent blocks: SelfAttention (model/attention.py:32-48), BaseLearner (model/capl.py:435-457)
# ----------------------------------------------------------------------------------------
def self_attention(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """eval mode (dropout off).  (B,256,N) -> (B,64,N)."""
    q = F.conv1d(x, sd[prefix + ".q_map.weight"])
    kk = F.conv1d(x, sd[prefix + ".k_map.weight"])
    v = F.conv1d(x, sd[prefix + ".v_map.weight"])
    temp = float(q.shape[1]) ** 0.5
    attn = torch.softmax(torch.matmul(q.transpose(1, 2) / temp, kk), dim=-1)
    return torch.matmul(attn, v.transpose(1, 2)).transpose(1, 2)


def base_learner(sd: SD, prefix: str, x: torch.Tensor, training: bool = False) -> torch.Tensor:
    i = 0
    n = 0
    while f"{prefix}.convs.{n}.0.weight" in sd:
        n += 1
    for i in range(n):
        x = F.conv1d(x, sd[f"{prefix}.convs.{i}.0.weight"], sd[f"{prefix}.convs.{i}.0.bias"])
        x = _bn(sd, f"{prefix}.convs.{i}.1", x, training)
        if i != n - 1:
            x = F.relu(x)
    return x


# ----------------------------------------------------------------------------------------
# GW head  (model/capl.py)
# ----------------------------------------------------------------------------------------
def get_features(sd: SD, gp: torch.Tensor, x: torch.Tensor, k: int = 20,
                 idx_list=None, knn: str = "formula", training: bool = False):
    """model/capl.py:324-362.  Returns dict with point_feat, semantic_feat, one_hot_feat and
    the intermediates the parity tests look at."""
    ecs, lvl2, used = dgcnn_forward(sd, x, k, "encoder.", idx_list, knn, training)
    lvl3 = base_learner(sd, "base_learner", lvl2, training)
    att = self_attention(sd, "att_learner", lvl2)
    ec = torch.cat(ecs, dim=1)
    semantic = torch.cat([ecs[0], att, lvl3], dim=1)
    cos = torch.matmul(F.normalize(gp, p=2, dim=1).unsqueeze(0), F.normalize(ec, p=2, dim=1))
    cosine_feat = torch.softmax(10 * cos, dim=1)
    assignment = cosine_feat.argmax(dim=1)
    one_hot = F.one_hot(assignment, num_classes=gp.shape[0]).transpose(2, 1).float()
    z = torch.cat([cosine_feat, semantic], dim=1)
    z = F.conv1d(z, sd["fusion.0.weight"], sd["fusion.0.bias"])
    z = F.leaky_relu(_bn(sd, "fusion.1", z, training), 0.2)
    return dict(point_feat=z, semantic_feat=semantic, one_hot_feat=one_hot, edge_convs=ec,
                feat_level2=lvl2, att_feat=att, feat_level3=lvl3, cos=cos,
                cosine_feat=cosine_feat, assignment=assignment, idx=used)


def get_pred(x: torch.Tensor, proto: torch.Tensor, bg_proto: Optional[torch.Tensor] = None) -> torch.Tensor:
    """model/capl.py:290-322 -- 10 * cos(proto, x) ; proto (cls,c) or (b,cls,c); optional bg row first."""
    if proto.dim() == 3:
        if bg_proto is not None:
            proto = torch.cat([bg_proto.unsqueeze(0).expand(proto.shape[0], -1, -1), proto], dim=1)
        pn = F.normalize(proto, p=2, dim=-1)
    else:
        if bg_proto is not None:
            proto = torch.cat([bg_proto, proto], dim=0)
        pn = F.normalize(proto, p=2, dim=1).unsqueeze(0)
    return torch.matmul(pn, F.normalize(x, p=2, dim=1)) * 10


def post_refine_proto_v2(proto: torch.Tensor, point_feat: torch.Tensor,
                         bg_proto: Optional[torch.Tensor] = None) -> torch.Tensor:
    """model/capl.py:245-287 -- query-adaptive prototype refinement -> (b, classes, c)."""
    pred = torch.softmax(get_pred(point_feat, proto, bg_proto), dim=2)
    pp = torch.matmul(pred, point_feat.transpose(1, 2))
    if bg_proto is not None:
        pp = pp[:, 1:, :]
    w = (F.normalize(pp, p=2, dim=-1) * F.normalize(proto, p=2, dim=-1).unsqueeze(0)).sum(-1, keepdim=True)
    w = w * (w > 0).float()
    return w * pp + (1 - w) * proto.unsqueeze(0)


def get_gp_weight(gp_coding: torch.Tensor, one_hot: torch.Tensor, th: float) -> torch.Tensor:
    """model/capl.py:92-142 (eval use: use_bg_weight=False) -- weight th where coding@one_hot == 1."""
    score = torch.matmul(gp_coding.unsqueeze(0).expand(one_hot.shape[0], -1, -1), one_hot)
    w = torch.ones_like(score)
    w[score == 1] = th
    return w


def gp_accuracies(gp_coding, one_hot, gt_label, base_num):
    """model/capl.py:104-114 diagnostic accuracies of the eval branch."""
    score = torch.matmul(gp_coding.unsqueeze(0).expand(one_hot.shape[0], -1, -1), one_hot)
    gt = F.one_hot(gt_label, num_classes=score.shape[1]).transpose(2, 1)
    per_point = (gt * score).sum(dim=1)
    acc = per_point.mean()
    novel = gt_label > base_num - 1
    novel_acc = per_point[novel].mean() if novel.sum() > 0 else torch.zeros_like(acc)
    return acc, novel_acc


def forward_eval(sd: SD, gp: torch.Tensor, x: torch.Tensor, gened_proto: torch.Tensor,
                 base_class_coding: torch.Tensor, novel_class_coding: torch.Tensor,
                 base_num: int, eval_weight: float, k: int = 20, idx_list=None, knn: str = "formula"):
    """model/capl.py:167-192 (eval_model=True branch).  Returns (logits (b,cls,n), features dict)."""
    f = get_features(sd, gp, x, k, idx_list, knn)
    pf = f["point_feat"]
    if gened_proto.dim() == 3:
        gened_proto = gened_proto[0]
    rp = post_refine_proto_v2(sd["main_proto"], pf)
    rp = rp.clone()
    rp[:, :base_num] = rp[:, :base_num] + gened_proto[:base_num].unsqueeze(0)
    rp[:, base_num:] = rp[:, base_num:] * 0 + gened_proto[base_num:].unsqueeze(0)
    logits = get_pred(pf, rp)
    coding = torch.cat([base_class_coding, novel_class_coding], dim=0)
    logits = logits * get_gp_weight(coding, f["one_hot_feat"], eval_weight)
    f["refine_proto"] = rp
    return logits, f


def generate_fake_proto(x: torch.Tensor, y: torch.Tensor, main_proto: torch.Tensor, fake_novel: Sequence[int]) -> torch.Tensor:
    """model/capl.py:364-411 with the sampled fake-novel class ids given (the reference draws them with random.sample)."""
    ty = y.unsqueeze(1)
    proto = main_proto / (main_proto.norm(2, 1, True) + 1e-12)
    xn = x / (x.norm(2, 1, True) + 1e-12)
    for fn in fake_novel:
        m = (ty == fn).to(x.dtype)
        feat = (xn * m).sum(0).sum(-1) / (m.sum(0).sum(-1) + 1e-12)
        sel = torch.zeros(proto.shape[0], 1, dtype=x.dtype)
        sel[int(fn) - 1] = 1
        proto = proto * (1 - sel) + feat.unsqueeze(0) * sel
    return proto


def forward_train(sd: SD, gp: torch.Tensor, x: torch.Tensor, y: torch.Tensor, base_num: int, fake_novel: Sequence[int],
                  k: int = 20, idx_list=None, knn: str = "formula", ignore_index: int = 255):
    """model/capl.py:194-242 (training branch, attention dropout off): returns (pred labels, loss, features dict).
    sd may hold tensors with requires_grad=True: the loss is differentiable w.r.t. them (autograd is the gradient oracle)."""
    f = get_features(sd, gp, x, k, idx_list, knn, training=True)
    pf = f["point_feat"]
    half = x.shape[0] // 2
    main = sd["main_proto"]
    ori = generate_fake_proto(pf[half:], y[half:], main.clone(), fake_novel)
    bg = sd["bg_proto"]
    l1 = F.cross_entropy(get_pred(pf, ori, bg), y, ignore_index=ignore_index)
    rp = post_refine_proto_v2(main.clone(), pf, bg)
    rp2 = rp.clone()
    rp2[:, :base_num] = rp2[:, :base_num] + ori[:base_num].unsqueeze(0)
    rp2[:, base_num:] = rp2[:, base_num:] * 0 + ori[base_num:].unsqueeze(0)
    logits2 = get_pred(pf, rp2, bg)
    l2 = F.cross_entropy(logits2, y, ignore_index=ignore_index)
    return logits2.max(1)[1], 0.5 * l2 + 0.5 * l1, f


# ----------------------------------------------------------------------------------------
# GW basis builder  (get_basis.py)
# ----------------------------------------------------------------------------------------
def kmean_to_proto(feat: np.ndarray, labels: np.ndarray, num_cnt: int) -> np.ndarray:
    """get_basis.py:27-44 -- per-cluster mean (asserts non-empty)."""
    rows = []
    for c in range(num_cnt):
        m = labels == c
        assert m.sum() != 0, f"empty cluster {c}"
        rows.append(feat[m].mean(axis=0))
    return np.stack(rows, axis=0)


def svd_reconstruct(protos: np.ndarray, energy: float = 0.95) -> np.ndarray:
    """get_basis.py:50-71 -- rank-r reconstruction, r = first rank whose singular-value mass > 95 %."""
    u, s, vh = np.linalg.svd(protos.T, full_matrices=False)
    r = len(s) - 1
    for i in range(len(s)):
        if s[: i + 1].sum() > energy * s.sum():
            r = i
            break
    rec = u[:, : r + 1] @ np.diag(s[: r + 1]) @ vh[: r + 1, :]
    return rec.T


def lloyd_reference(X: np.ndarray, init: np.ndarray, max_iter: int = 300, tol: float = 1e-4):
    """sklearn 1.9.0 KMeans(init=<array>, n_init=1).fit as get_basis.py:210 drives it, restated with the
    pinned E-step: mean-centring (_kmeans.py:1487-1490), tol scaling (_kmeans.py:289-296), Lloyd loop and
    convergence (_kmeans.py:630-759).  Empty clusters raise (the caller asserts non-empty,
    get_basis.py:37).  Returns (labels int32, centers fp32 (un-centred), n_iter)."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    mean = X.mean(axis=0)
    Xc = X - mean
    centers = np.ascontiguousarray(init, dtype=np.float32) - mean
    K = centers.shape[0]
    tol_abs = float(np.mean(np.var(Xc, axis=0)) * tol)
    labels_old = np.full(X.shape[0], -1, dtype=np.int32)
    strict = False
    it = 0
    for it in range(1, max_iter + 1):
        labels = kmeans_assign_exact(Xc, centers)
        sums, counts = kmeans_accumulate_exact(Xc, labels, K)
        empty = np.where(counts == 0)[0]
        if len(empty):
            # sklearn _k_means_common.pyx:167-211: move each empty centre onto one of the points
            # farthest from its own (old) centre, taking that point out of its old cluster's sum.
            dist = ((Xc - centers[labels]) ** 2).sum(axis=1)
            if dist.max() > 0:
                far = np.argpartition(dist, -len(empty))[: -len(empty) - 1: -1]
                for e, fi in zip(empty, far):
                    old = labels[fi]
                    sums[old] -= Xc[fi]
                    sums[e] = Xc[fi]
                    counts[e] = 1
                    counts[old] -= 1
        new = np.empty_like(centers)
        big = int(np.argmax(counts))
        for c in range(K):  # _k_means_common.pyx:236-260 (_average_centers)
            new[c] = (sums[c] / counts[c]).astype(np.float32) if counts[c] > 0 else 0
        for c in range(K):
            if counts[c] <= 0:
                new[c] = new[big]
        shift = float(((new - centers).astype(np.float64) ** 2).sum())
        centers = new
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if shift <= tol_abs:
            break
        labels_old = labels
    if not strict:
        labels = kmeans_assign_exact(Xc, centers)
    return labels, centers + mean, it


# ----------------------------------------------------------------------------------------
# synthetic inputs -- shared by tests, smoke and bench
# ----------------------------------------------------------------------------------------
def synthetic_blocks(B: int, N: int, seed: int = 1234, dup_frac: float = 0.0) -> torch.Tensor:
    """(B, 9, N) fp32 S3DIS-shaped blocks as dataloaders/loader.py:87-101 emits them:
    ch0-2 xyz (x,y in [0,1], z in [0,3]), ch3-5 rgb in [0,1], ch6-8 XYZ normalised per block.
    dup_frac > 0 makes that fraction of points exact copies (sampling with replacement, loader.py:66)."""
    out = torch.empty(B, 9, N, dtype=torch.float32)
    for b in range(B):
        g = torch.Generator().manual_seed(seed + b)
        xyz = torch.rand(3, N, generator=g) * torch.tensor([[1.0], [1.0], [3.0]])
        rgb = torch.rand(3, N, generator=g)
        if dup_frac > 0:
            nd = int(N * dup_frac)
            src = torch.randint(0, N, (nd,), generator=g)
            dst = torch.randperm(N, generator=g)[:nd]
            xyz[:, dst] = xyz[:, src]
            rgb[:, dst] = rgb[:, src]
        xyz = xyz - xyz.min(dim=1, keepdim=True).values
        out[b, 0:3] = xyz
        out[b, 3:6] = rgb
        out[b, 6:9] = xyz / xyz.max(dim=1, keepdim=True).values.clamp_min(1e-12)
    return out


def randomize_bn_(sd: SD, seed: int = 5) -> SD:
    """Make BN folding non-trivial: gamma~N(1,.3) beta~N(0,.2) mean~N(0,.2) var~U[.5,2]."""
    g = torch.Generator().manual_seed(seed)
    for key in sorted(sd.keys()):
        if key.endswith("running_mean"):
            p = key[: -len("running_mean")]
            n = sd[key].numel()
            sd[p + "weight"] = 1.0 + 0.3 * torch.randn(n, generator=g)
            sd[p + "bias"] = 0.2 * torch.randn(n, generator=g)
            sd[p + "running_mean"] = 0.2 * torch.randn(n, generator=g)
            sd[p + "running_var"] = 0.5 + 1.5 * torch.rand(n, generator=g)
    return sd


# ---------------------------------------------------------------------------------------------------------------
# N4 / N2 callers of the hot path: the mIoU metric and the geometric-word class codings
# ---------------------------------------------------------------------------------------------------------------
def evaluate_metric(pred_labels_list, gt_labels_list, test_classes, novel_classes, all_learning_order, scannet=False):
    """runs/eval.py:10-108 (evaluate_metric_GFS) without the logger: the per-point Python loop of :31-48 restated with
    numpy bincounts (integer counts, any order), the IoU arithmetic of :50-106 kept operation for operation.
    -> (mean_iou, base_iou, novel_iou, hm, iou array)"""
    num_class = len(test_classes)
    order = np.asarray(all_learning_order, dtype=np.int64)
    gt_classes = np.zeros(num_class, dtype=np.int64)
    positive = np.zeros(num_class, dtype=np.int64)
    true_pos = np.zeros(num_class, dtype=np.int64)
    for pred, gt in zip(pred_labels_list, gt_labels_list):
        p = np.asarray(pred).reshape(-1).astype(np.int64)
        g = np.asarray(gt).reshape(-1).astype(np.int64)
        gt_classes += np.bincount(order[g], minlength=num_class)                 # :40-41
        positive += np.bincount(order[p], minlength=num_class)                   # :44-45
        true_pos += np.bincount(order[g][g == p], minlength=num_class)           # :47
    iou_list, base, novel = [], [], []
    for c in range(num_class):
        iou = int(true_pos[c]) / float(int(gt_classes[c]) + int(positive[c]) - int(true_pos[c]))
        iou_list.append(iou)
        if scannet and c == 0:
            continue
        (novel if c in novel_classes else base).append(iou)
    if scannet:
        mean_iou = np.array(iou_list[1:]).mean()
        iou_list = iou_list[1:]
    else:
        mean_iou = np.array(iou_list).mean()
    base_iou, novel_iou = np.array(base).mean(), np.array(novel).mean()
    hm = 2 * base_iou * novel_iou / (base_iou + novel_iou)
    return mean_iou, base_iou, novel_iou, hm, np.array(iou_list)


def hard_coding(coding, energy):
    """train.py:136-152 (post_processing_hard_coding): multi-hot of the most frequent words holding > energy of the mass.
    Sequential fp32 running sum; equal frequencies lowest index first (torch.argsort leaves that open)."""
    c = np.asarray(coding, dtype=np.float32).copy()
    total = torch.sum(torch.from_numpy(c))
    thr = float((energy * total).item())
    acc = np.float32(0)
    mask = np.zeros(c.shape, dtype=bool)
    for i in np.argsort(-c, kind="stable"):
        acc = np.float32(acc + c[i])
        mask[i] = True
        if float(acc) > thr:
            break
    return mask.astype(np.float32)


def class_gw_codings(assignments, labels, train_class, G, energy):
    """train.py:156-218 (collect_base_class_gp_coding_sum) given the per-block GW assignments instead of the model:
    assignments / labels: lists of (n,) integer arrays, one per block (label 0 = background, cls + 1 = base class cls).
    -> (base_class_gp_coding (num_base, G) float32 multi-hot, bg_class_coding (G,) float32, freq (num_base, G) float32 = the
    word frequencies the multi-hot was cut from); at most 2000 blocks: no sampling"""
    sums = {cls: np.zeros(G, dtype=np.float32) for cls in train_class}
    counts = {cls: 0 for cls in train_class}
    bg = []
    for a, t in zip(assignments, labels):
        a = np.asarray(a).astype(np.int64)
        t = np.asarray(t).astype(np.int64)
        for cls in np.unique(t):
            m = t == cls
            h = np.bincount(a[m], minlength=G).astype(np.float32)
            if cls == 0:
                bg.append(h / np.float32(m.sum()))                              # :187-190  mean of the one-hot columns
            elif cls - 1 in sums:
                sums[cls - 1] += h                                              # :196-199
                counts[cls - 1] += int(m.sum())
    freq = np.stack([sums[cls] / np.float32(counts[cls]) for cls in train_class], axis=0)
    coding = np.stack([hard_coding(f, energy) for f in freq], axis=0)
    assert len(bg) <= 2000, "the reference samples 2000 blocks at random beyond that (train.py:213-214)"
    bg_coding = torch.mean(torch.from_numpy(np.stack(bg, axis=0)), dim=0).numpy()
    return coding, bg_coding, freq


def codings_equal_modulo_ties(freq, a, b):
    """two multi-hot codings cut from the same frequencies agree if they keep the same NUMBER of words, every word more
    frequent than the least frequent kept one, and none less frequent: `torch.argsort` orders equal frequencies arbitrarily
    (train.py:143), so which of several equally frequent words crosses the energy threshold is not defined by the reference"""
    freq, a, b = np.asarray(freq), np.asarray(a) > 0.5, np.asarray(b) > 0.5
    if a.shape != b.shape or a.shape != freq.shape:
        return False
    for f, x, y in zip(freq.reshape(-1, freq.shape[-1]), a.reshape(-1, a.shape[-1]), b.reshape(-1, b.shape[-1])):
        if x.sum() != y.sum():
            return False
        cut = f[x].min()
        if not (np.array_equal(x[f > cut], y[f > cut]) and x[f > cut].all() and not x[f < cut].any() and not y[f < cut].any()):
            return False
    return True


def kmeans_plusplus_ref(X, n_clusters, rng):
    """sklearn 1.9.0 _kmeans.py:_kmeans_plusplus restated in numpy: the random stream is consumed exactly as sklearn consumes
    it (choice(n, p=w/sum(w)) for the first centre, then uniform(size=2+ln k) per centre), the distances are the float64
    evaluation rounded to float32 and clamped at 0 (metrics/pairwise.py:_euclidean_distances_upcast).  The two places where
    sklearn itself is not reproducible by a parallel implementation -- the potential (a float32 BLAS dot) and the cumulative
    sum (np.cumsum in float32) -- are float64 here, as in gfs3d.kmeans; tests/golden/kmeanspp_*.npz hold sklearn's own picks
    and pin this restatement against them.  X: the float32 matrix k-means++ runs on (KMeans.fit passes the mean-centred data).
    -> (centres (k, D) float32, indices (k,))"""
    X = np.asarray(X, dtype=np.float32)
    n = X.shape[0]
    trials = 2 + int(np.log(n_clusters))
    X64 = X.astype(np.float64)
    xsq = (X64 * X64).sum(1)

    def dist(c):
        return np.maximum((xsq[None, :] - 2.0 * (X64[c] @ X64.T) + xsq[c][:, None]).astype(np.float32), 0)

    idx = np.empty(n_clusters, dtype=np.int64)
    w = np.ones(n, dtype=np.float32)
    idx[0] = int(rng.choice(n, p=w / w.sum()))
    closest = dist(idx[:1])[0]
    pot = closest.astype(np.float64).sum()
    for c in range(1, n_clusters):
        rv = rng.uniform(size=trials) * pot
        cand = np.minimum(np.searchsorted(np.cumsum(closest.astype(np.float64)), rv), n - 1)
        d = np.minimum(dist(cand), closest[None, :])
        pots = d.astype(np.float64).sum(1)
        best = int(np.argmin(pots))
        pot, closest, idx[c] = pots[best], d[best], cand[best]
    return X[idx], idx
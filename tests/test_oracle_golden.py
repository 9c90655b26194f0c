"""CPU suite: the oracle restatement against the fixtures generated from the real reference
(tests/golden/make_golden.py).  This is the pin the GPU parity tests rest on."""
import numpy as np
import pytest
import torch

from oracle import gfs_oracle as O
from parity import knn_classify_mismatches, rel_err

torch.set_num_threads(max(1, torch.get_num_threads()))


@pytest.mark.parametrize("name", ["dgcnn_b2_n256", "dgcnn_dup_b1_n256", "dgcnn_b1_n320_k40", "dgcnn_b1_n2048"])
def test_dgcnn_oracle_matches_reference(golden, golden_sd, name):
    g = golden(name)
    sd = golden_sd("dgcnn_weights")
    x = torch.from_numpy(g["x"])
    k, s = int(g["k"]), int(g["subsample"])
    with torch.no_grad():
        ecs, out, idx = O.dgcnn_forward(sd, x, k, "", knn="formula")
    # same ATen ops in the same order as model/dgcnn.py:113-127 -> bit-identical on the same torch build;
    # 1e-6 leaves room for a different CPU's BLAS kernel selection
    assert rel_err(torch.cat(ecs, 1)[:, :, ::s], torch.from_numpy(g["ec"])) <= 1e-6
    assert rel_err(out[:, :, ::s], torch.from_numpy(g["out"])) <= 1e-6


@pytest.mark.parametrize("name", ["dgcnn_b2_n256", "dgcnn_dup_b1_n256", "dgcnn_b1_n320_k40", "dgcnn_b1_n2048"])
def test_knn_exact_vs_reference_indices(golden, name):
    """pinned-order kNN (C) vs the reference's own topk indices: neighbour sets equal except fp32 near-ties."""
    g = golden(name)
    x = torch.from_numpy(g["x"])
    k = int(g["k"])
    idx, dist = O.knn_exact(x, k, return_dist=True)
    n, near, real = knn_classify_mismatches(x, idx.numpy(), g["idx0"].astype(np.int64), k)
    assert real == 0, f"{real} rows differ beyond fp32 rounding ({n} mismatching rows, {near} near-ties)"
    if "dup" not in name:   # exact duplicates make exact ties common: which copy wins is undefined in the reference
        assert n <= 0.002 * x.shape[0] * x.shape[2] + 2
    # structure: sorted nearest-first, ties ascending index, self (distance 0) is present
    d = dist.numpy()
    assert (d[..., :-1] >= d[..., 1:]).all()
    tie = d[..., :-1] == d[..., 1:]
    ii = idx.numpy()
    assert (ii[..., :-1][tie] < ii[..., 1:][tie]).all()
    self_in = (ii == np.arange(x.shape[2])[None, :, None]).any(-1)
    assert self_in.all()


def test_knn_exact_duplicates_are_distinct_neighbours(golden):
    g = golden("dgcnn_dup_b1_n256")
    x = torch.from_numpy(g["x"])
    idx, dist = O.knn_exact(x, 20, return_dist=True)
    # exact copies give distance exactly 0 and are returned as separate neighbours (loader.py:66 replace=True)
    assert ((dist.numpy() == 0).sum(-1) >= 1).all()
    assert ((dist.numpy() == 0).sum(-1) >= 2).any()
    for row in idx.numpy().reshape(-1, 20)[:64]:
        assert len(set(row.tolist())) == 20


@pytest.mark.parametrize("name,wname", [("gfs_s3dis_b2_n256", "gfs_s3dis_weights"),
                                        ("gfs_scannet_b2_n128", "gfs_scannet_weights")])
def test_gfs_eval_oracle_matches_reference(golden, golden_sd, name, wname):
    g = golden(name)
    sd = golden_sd(wname)
    t = lambda k: torch.from_numpy(g[k])
    with torch.no_grad():
        logits, f = O.forward_eval(sd, t("gp"), t("x"), t("gened_proto"), t("base_class_coding"),
                                   t("novel_class_coding"), int(g["base_num"]), float(g["eval_weight"]))
    assert rel_err(f["point_feat"], t("point_feat")) <= 1e-6
    assert rel_err(f["semantic_feat"], t("semantic_feat")) <= 1e-6
    assert rel_err(logits, t("logits")) <= 1e-6
    assert (f["assignment"].numpy() == g["assignment"]).mean() >= 0.999
    assert (logits.argmax(1).numpy() == g["logits"].argmax(1)).mean() >= 0.999
    acc, nacc = O.gp_accuracies(torch.cat([t("base_class_coding"), t("novel_class_coding")]), f["one_hot_feat"],
                                t("y").long(), int(g["base_num"]))
    assert abs(float(acc) - float(g["gp_acc"])) < 1e-6 and abs(float(nacc) - float(g["gp_novel_acc"])) < 1e-6


def _kmeans_X(g):
    rs = np.random.RandomState(int(g["seed"]))
    n, D, K = int(g["n"]), int(g["D"]), int(g["K"])
    cent = rs.randn(K, D).astype(np.float32)
    lab = rs.randint(0, K, size=n)
    X = (cent[lab] + 0.35 * rs.randn(n, D)).astype(np.float32)
    assert abs(X.astype(np.float64).sum() - float(g["x_checksum"])) < 1e-6
    return X


@pytest.mark.parametrize("name", ["kmeans_n6000_k150", "kmeans_n2000_k20"])
def test_kmeans_oracle_matches_sklearn_fixture(golden, name):
    g = golden(name)
    X = _kmeans_X(g)
    labels, centers, n_iter = O.lloyd_reference(X, g["init"])
    assert (labels == g["labels"]).mean() >= 0.999          # sklearn's sgemm order differs: near-ties only
    assert np.abs(centers - g["centers"]).max() <= 1e-5
    assert n_iter == int(g["n_iter"])
    basis = O.svd_reconstruct(O.kmean_to_proto(X, g["labels"].astype(np.int64), int(g["K"])))
    assert np.abs(basis - g["basis"]).max() <= 1e-5


@pytest.mark.parametrize("name", ["kmeanspp_n3000_k50", "kmeanspp_n6000_k150"])
def test_kmeans_plusplus_oracle_reproduces_sklearns_picks(golden, name):
    """the k-means++ restatement (sklearn's random-stream order, float64-evaluated distances rounded to float32) picks the
    centres sklearn.cluster.kmeans_plusplus picked for the same RandomState (fixture written by make_golden.py:make_kmeanspp)"""
    g = golden(name)
    n, D, K = int(g["n"]), int(g["D"]), int(g["K"])
    rs = np.random.RandomState(int(g["seed"]))
    cent = rs.randn(K, D).astype(np.float32)
    X = (cent[rs.randint(0, K, n)] + 0.35 * rs.randn(n, D)).astype(np.float32)
    X = X - X.mean(0)
    assert abs(float(X.astype(np.float64).sum()) - float(g["x_checksum"])) < 1e-9
    _, idx = O.kmeans_plusplus_ref(X, K, np.random.RandomState(int(g["rs_seed"])))
    assert np.array_equal(idx, g["idx"])


def test_kmeans_assign_ties_lowest_index():
    X = np.zeros((4, 8), np.float32)
    C = np.zeros((5, 8), np.float32)      # all centroids identical -> label 0 (strict <, _k_means_lloyd.pyx:208)
    assert (O.kmeans_assign_exact(X, C) == 0).all()


@pytest.mark.parametrize("name", ["metric_s3dis", "metric_scannet"])
def test_metric_oracle_matches_reference_eval(golden, name):
    """oracle.evaluate_metric == the reference's runs/eval.py (evaluate_metric_GFS) on the committed labels, float for float"""
    g = golden(name)
    ncls = len(g["order"])
    got = O.evaluate_metric(list(g["pred"]), list(g["gt"]), list(range(ncls)), g["novel"].tolist(), g["order"].tolist(),
                            scannet=bool(g["scannet"]))
    assert got[0] == float(g["mean_iou"]) and got[1] == float(g["base_iou"])
    assert got[2] == float(g["novel_iou"]) and got[3] == float(g["hm"])
    assert np.array_equal(got[4], g["ious"])


def test_class_coding_oracle_matches_reference_train_functions(golden):
    """oracle.class_gw_codings == train.py:156-218 run on a stub model's one-hot GW features (modulo the order torch.argsort
    gives to equally frequent words); the background coding is a plain mean and must agree to the last bit"""
    g = golden("coding_s3dis")
    nb, G = int(g["num_base"]), int(g["G"])
    coding, bg, freq = O.class_gw_codings(list(g["assign"]), list(g["labels"]), list(range(nb)), G, float(g["energy"]))
    assert np.array_equal(freq, g["freq"])
    assert O.codings_equal_modulo_ties(freq, coding, g["coding"])
    assert np.array_equal(bg, g["bg_coding"])
    assert coding.sum(1).tolist() == g["coding"].sum(1).tolist()


def test_hard_coding_keeps_the_smallest_prefix_above_the_energy():
    c = np.array([0.05, 0.4, 0.05, 0.3, 0.2], dtype=np.float32)
    assert O.hard_coding(c, 0.85).tolist() == [0, 1, 0, 1, 1]          # 0.4 + 0.3 = 0.7 <= 0.85 < 0.9
    assert O.hard_coding(c, 0.95).tolist() == [1, 1, 0, 1, 1]          # equal frequencies: lowest index first
    assert O.hard_coding(c, 1.0).tolist() == [1, 1, 1, 1, 1]           # never strictly above the total: keep everything

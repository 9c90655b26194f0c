#!/usr/bin/env python
"""Measure the two callers widened in round 1 (N4 metric, N2 class codings) on one GPU, with the
reference's own CPU procedure timed beside them on a bounded sample.  One JSON line per caller."""
import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gfs-3dseg_gws_b200"))
import bench  # noqa: E402
from gfs3d import ops  # noqa: E402
from gfs3d.coding import collect_base_class_gp_coding_sum  # noqa: E402
from gfs3d.synthetic import synthetic_blocks  # noqa: E402
from runs.eval import evaluate_metric_GFS  # noqa: E402


def ev_ms(fn, iters=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def reference_metric_loop(pred, gt, order, ncls):
    """runs/eval.py:31-48 as the reference runs it: one Python iteration per point"""
    gt_c, pos_c, tp_c = [0] * ncls, [0] * ncls, [0] * ncls
    for j in range(pred.shape[0]):
        for k in range(pred.shape[1]):
            g, p = int(gt[j, k]), int(pred[j, k])
            gt_c[order[g]] += 1
            pos_c[order[p]] += 1
            tp_c[order[g]] += int(g == p)
    return gt_c, pos_c, tp_c


def main():
    dev = torch.device("cuda", 0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))

    # ---- N4: mIoU metric over a whole evaluation (32768 blocks x 2048 points = 64 M points)
    ncls, n = 13, 64 << 20
    g = torch.Generator(device="cuda").manual_seed(1)
    gt = torch.randint(0, ncls, (n // 2048, 2048), device=dev, dtype=torch.int32, generator=g)
    pred = torch.where(torch.rand(gt.shape, device=dev, generator=g) < 0.7, gt, torch.randint(0, ncls, gt.shape, device=dev, dtype=torch.int32, generator=g))
    ms = ev_ms(lambda: ops.joint_histogram(gt, pred, ncls, ncls))
    log = SimpleNamespace(cprint=lambda *_: None)
    order = list(range(ncls))
    t0 = time.perf_counter()
    out = evaluate_metric_GFS(log, [pred], [gt], list(range(ncls)), [10, 11, 12], order)
    torch.cuda.synchronize()
    t_api = time.perf_counter() - t0
    sample = 200_000
    ps, gs = pred.reshape(-1)[:sample].cpu().numpy().reshape(-1, 2048 if sample % 2048 == 0 else sample), None
    ps = pred.reshape(-1)[:sample].cpu().numpy().reshape(1, -1)
    gs = gt.reshape(-1)[:sample].cpu().numpy().reshape(1, -1)
    t0 = time.perf_counter()
    reference_metric_loop(ps, gs, order, ncls)
    t_ref = time.perf_counter() - t0
    print(json.dumps({"caller": "evaluate_metric_GFS (runs/eval.py:31-48)", "points": n, "kernel_ms": ms,
                      "kernel_points_per_s": n / (ms / 1e3), "algorithmic_gbs": n * 8 / 1e9 / (ms / 1e3), "hbm_peak_gbs": hbm,
                      "frac_of_hbm_peak": n * 8 / 1e9 / (ms / 1e3) / hbm, "api_seconds_incl_host_math": t_api,
                      "mean_iou": float(out[0]),
                      "cpu_reference": {"points_per_s": sample / t_ref, "cores": 1, "kind": "port", "sample": f"{sample} points, per-point Python loop"}}))

    # ---- N2: base-class GW codings over a training set of 1024 blocks
    m, gp = bench.build_model(dev)
    nblk, B = 1024, 32
    xs = [synthetic_blocks(B, bench.NPTS, seed=50 + i).to(dev) for i in range(4)]
    gl = torch.Generator().manual_seed(3)
    ys = [torch.randint(0, bench.BASE_NUM + 1, (B, bench.NPTS), generator=gl).to(dev) for _ in range(4)]
    loader = [(xs[i % 4], ys[i % 4], None) for i in range(nblk // B)]
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        collect_base_class_gp_coding_sum(m, loader[:2], list(range(bench.BASE_NUM)), 0.9)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        coding, bg = collect_base_class_gp_coding_sum(m, loader, list(range(bench.BASE_NUM)), 0.9)
        torch.cuda.synchronize()
        t_ours = time.perf_counter() - t0

        # the reference's procedure (train.py:156-218): batch size 1, one-hot features, per-class masks with host syncs
        def ref_style(blocks):
            feats = {c: [] for c in range(bench.BASE_NUM)}
            nums = {c: [] for c in range(bench.BASE_NUM)}
            with torch.no_grad():
                for i in range(blocks):
                    x1, t1 = xs[i % 4][i % B:i % B + 1], ys[i % 4][i % B]
                    _, _, gp_feat = m.getFeatures(x1)
                    gp_feat = gp_feat[0]
                    for cls in torch.unique(t1):
                        mask = t1 == cls
                        if cls == 0:
                            torch.mean(gp_feat[:, mask], dim=1)
                            continue
                        if torch.sum(mask) > 0:
                            feats[cls.item() - 1].append(torch.sum(gp_feat[:, mask], dim=1))
                            nums[cls.item() - 1].append(torch.sum(mask))
            torch.cuda.synchronize()
        ref_style(4)
        t0 = time.perf_counter()
        ref_blocks = 64
        ref_style(ref_blocks)
        t_ref = time.perf_counter() - t0
    print(json.dumps({"caller": "collect_base_class_gp_coding_sum (train.py:156-218)", "blocks": nblk, "batch": B,
                      "seconds": t_ours, "blocks_per_s": nblk / t_ours, "words_kept_per_class": coding.sum(1).int().tolist(),
                      "reference_procedure_on_this_gpu": {"blocks_per_s": ref_blocks / t_ref, "sample": f"{ref_blocks} blocks, batch size 1, "
                                                          "one-hot features and per-class masks as in the reference, same fused model underneath"}}))


if __name__ == "__main__":
    main()

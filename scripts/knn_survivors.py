"""Diagnostics: how selective is the tensor-core kNN filter on the three graphs of the bench model (random-init weights)?
Prints, per kNN call of one eval step, the share of repaired 64-row tiles and basic statistics of the features."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gfs-3dseg_gws_b200"))
import bench  # noqa: E402
from gfs3d import ops  # noqa: E402
from gfs3d.synthetic import synthetic_blocks  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    m, gp = bench.build_model(dev)
    gened, bc, nc = bench.head_inputs(dev)
    x = synthetic_blocks(8, bench.NPTS, seed=1234).to(dev)
    seen = []
    real = ops.knn

    def spy(t, k, *a, **kw):
        seen.append(t.clone())
        return real(t, k, *a, **kw)

    ops.knn = spy
    import model.dgcnn as dg
    if hasattr(dg, "ops"):
        dg.ops.knn = spy
    with torch.no_grad():
        m(x=x, y=None, eval_model=True, gened_proto=gened, base_class_coding=bc, novel_class_coding=nc)
    ops.knn = real
    for i, t in enumerate(seen):
        t = t.contiguous()
        idx, filt, flags = ops.knn_tc_diag(t, 20)
        torch.cuda.synchronize()
        N4 = t.shape[2] // 4
        mu = 0.25 * ((t[:, :, 0] + t[:, :, N4]) + (t[:, :, 2 * N4] + t[:, :, 3 * N4]))
        xc = t - mu[:, :, None]
        cc = (xc * xc).sum(1)
        xx = (t * t).sum(1)
        B, C, N = t.shape
        a = 2.0 ** -15 * cc + (C + 4) * 2.0 ** -25 * xx
        # exact distance gap between the 20th and the 21st neighbour, in units of the row's margin 2 a_i
        d = ops.knn(t, 20, return_dist=True, impl="exact")[1]
        print(f"kNN call {i}: C={C} repaired tiles {int(flags.sum())}/{flags.numel()}  mean |x~|^2 {float(cc.mean()):.3g} "
              f"mean |x|^2 {float(xx.mean()):.3g}  mean margin {float(2 * a.mean()):.3g}  "
              f"mean |d_k| {float(d[..., -1].abs().mean()):.3g}  mean (d_1 - d_k) {float((d[..., 1] - d[..., -1]).mean()):.3g}")


if __name__ == "__main__":
    main()

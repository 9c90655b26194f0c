"""kNN development probe: per-call timings (exact / tc ordered / tc set) and survivor statistics of the tensor-core filter on
the three graphs of the bench model (batch 32 x 2048 points, random-init weights).  One JSON line per kNN call."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gfs-3dseg_gws_b200"))
import bench  # noqa: E402
from gfs3d import ops  # noqa: E402
from gfs3d._lib import lib  # noqa: E402
from gfs3d.synthetic import synthetic_blocks  # noqa: E402

KT_CAP = 128


def time_ms(fn, iters=10, warm=2):
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


def survivor_counts(t, k):
    """run gfs_knn_tc_f32 on our own workspace and read the per-row survivor counts back (layout of kt_plan in knn_tc.cu)"""
    B, C, N = t.shape
    npad = (N + 255) // 256 * 256
    cp16 = (C + 15) // 16 * 16
    cpt = 16 if C <= 16 else 64
    kb = (2 * cp16 + 63) // 64
    o = B * (npad // 128) * kb * 16384 + B * npad * cpt * 4 + 3 * B * npad * 4
    off_cnt = o + B * N * 4 * KT_CAP * 4
    nbytes = int(lib().gfs_knn_tc_workspace_bytes(B, C, N))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=t.device)
    sq = torch.empty(B, N, device=t.device)
    idx = torch.empty(B, N, k, dtype=torch.int32, device=t.device)
    rc = lib().gfs_knn_tc_f32(t.data_ptr(), t.stride(0), B, C, N, k, sq.data_ptr(), ws.data_ptr(), nbytes, idx.data_ptr(), None,
                              torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    cnt = ws[off_cnt:off_cnt + B * N * 8].view(torch.int32).view(B * N, 2).long()
    return cnt


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    m, gp = bench.build_model(dev)
    gened, bc, nc = bench.head_inputs(dev)
    x = synthetic_blocks(32, bench.NPTS, seed=1234).to(dev)
    seen = []
    real = ops.knn

    def spy(t, k, *a, **kw):
        seen.append(t.clone())
        return real(t, k, *a, **kw)

    ops.knn = spy
    with torch.no_grad():
        m(x=x, y=None, eval_model=True, gened_proto=gened, base_class_coding=bc, novel_class_coding=nc)
    ops.knn = real
    for i, t in enumerate(seen):
        t = t.contiguous()
        k = 20
        a = ops.knn(t, k, impl="exact")
        b = ops.knn(t, k, impl="tc")
        c = ops.knn(t, k, impl="tc", ordered=False)
        cnt = survivor_counts(t, k)
        tot = cnt.sum(1).float()
        line = {"call": i, "C": t.shape[1], "identical_ordered": bool(torch.equal(a, b)),
                "identical_set": bool(torch.equal(a.sort(-1).values, c.sort(-1).values)),
                "survivors_mean": float(tot.mean()), "survivors_p99": float(tot.quantile(0.99)), "survivors_max": float(tot.max()),
                "flagged_rows": int((cnt < 0).any(1).sum()),
                "ms_exact": time_ms(lambda: ops.knn(t, k, impl="exact")),
                "ms_tc_ordered": time_ms(lambda: ops.knn(t, k, impl="tc")),
                "ms_tc_set": time_ms(lambda: ops.knn(t, k, impl="tc", ordered=False))}
        print(json.dumps(line), flush=True)
    # k = 40 (config[4])
    t = seen[1].contiguous()
    a = ops.knn(t, 40, impl="exact")
    c = ops.knn(t, 40, impl="tc", ordered=False)
    print(json.dumps({"k": 40, "C": 64, "identical_set": bool(torch.equal(a.sort(-1).values, c.sort(-1).values)),
                      "ms_exact": time_ms(lambda: ops.knn(t, 40, impl="exact")),
                      "ms_tc_set": time_ms(lambda: ops.knn(t, 40, impl="tc", ordered=False))}), flush=True)


if __name__ == "__main__":
    main()

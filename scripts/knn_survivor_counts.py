"""Diagnostics: survivor records per query row that the tensor-core kNN filter hands to the finish kernel, for the three
graphs of the bench model (random-init weights).  Reads the counters out of the call's workspace (one chain)."""
import ctypes
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


def plan(B, C, N):
    npad = (N + 255) // 256 * 256
    cp16 = (C + 15) // 16 * 16
    cpt = 16 if C <= 16 else 64
    kb = (2 * cp16 + 63) // 64
    o = B * (npad // 128) * kb * 16384
    o += B * npad * cpt * 4 + 3 * B * npad * 4
    off_surv = o
    o += B * N * 4 * KT_CAP * 4
    return off_surv, o


def main():
    dev = torch.device("cuda", 0)
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
    lib().gfs_knn_tc_set_split(0)
    k = 20
    for i, t in enumerate(seen):
        t = t.contiguous()
        B, C, N = t.shape
        nbytes = int(lib().gfs_knn_tc_workspace_bytes(B, C, N))
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        sq = torch.empty(B, N, device=dev)
        idx = torch.empty(B, N, k, dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        rc = lib().gfs_knn_tc_set_f32(ctypes.c_void_p(t.data_ptr()), t.stride(0), B, C, N, k, ctypes.c_void_p(sq.data_ptr()),
                                      ctypes.c_void_p(ws.data_ptr()), nbytes, ctypes.c_void_p(idx.data_ptr()), ctypes.c_void_p(st))
        assert rc == 0
        torch.cuda.synchronize()
        off_surv, off_cnt = plan(B, C, N)
        cnt = ws[off_cnt:off_cnt + B * N * 8].view(torch.int32).view(B * N, 2)
        ns = cnt.sum(1).float()
        ok = (cnt >= 0).all(1)
        fast = ok & (ns <= 64) & (ns - k <= 12)
        q = torch.quantile(ns[ok], torch.tensor([0.5, 0.9, 0.99], device=dev))
        ts = []
        for mode in ("set", "ordered"):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.knn(t, k, impl="tc", ordered=(mode == "ordered"))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 5)
        print(f"kNN call {i}: C={C}  survivors/row mean {float(ns[ok].mean()):.1f} median {float(q[0]):.0f} p90 {float(q[1]):.0f} "
              f"p99 {float(q[2]):.0f} max {float(ns[ok].max()):.0f}; rows on the bound-classification path {float(fast.float().mean()):.3f}; "
              f"flagged rows {int((~ok).sum())}; call (one chain) set {ts[0] * 1e3:.0f} us, ordered {ts[1] * 1e3:.0f} us")


if __name__ == "__main__":
    main()

"""Time the two kNN kernels (all-fp32 vs tensor-core filtered) on the bench shapes; prints one JSON line per case."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gfs-3dseg_gws_b200"))
from gfs3d import ops  # noqa: E402


def time_ms(fn, iters=20, warm=3):
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


def main():
    B, N, k = 32, 2048, 20
    g = torch.Generator().manual_seed(0)
    for C in (9, 64):
        if C == 9:
            x = torch.rand(B, C, N, generator=g).cuda()
        else:
            pos = torch.rand(B, 3, N, generator=g)
            w = torch.randn(64, 3, generator=g)
            x = torch.nn.functional.leaky_relu(torch.einsum("oc,bcn->bon", w, pos) + 0.1 * torch.randn(B, 64, N, generator=g), 0.2).cuda().contiguous()
        out = {}
        for impl in ("exact", "tc"):
            out[impl] = time_ms(lambda: ops.knn(x, k, impl=impl))
        a = ops.knn(x, k, impl="exact")
        b = ops.knn(x, k, impl="tc")
        _, _, flags = ops.knn_tc_diag(x[:4].contiguous(), k)
        print(json.dumps({"C": C, "B": B, "N": N, "k": k, "ms_exact": out["exact"], "ms_tc": out["tc"],
                          "identical": bool(torch.equal(a, b)), "repaired_tiles_of_first_4_blocks": int(flags.sum())}))


if __name__ == "__main__":
    main()

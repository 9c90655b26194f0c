"""rowsel development probe: GW projection (bench shape) and k-means E-step (1 M x 192, 150 centres), fp32 vs tensor-core path"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gfs-3dseg_gws_b200"))
from gfs3d import ops

def time_ms(fn, iters=10, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

torch.cuda.set_device(0)
g = torch.Generator(device="cuda").manual_seed(0)
B, N, G = 32, 2048, 150
ec = torch.randn(B, 192, N, device="cuda", generator=g).abs() * 0.3
gp = torch.nn.functional.normalize(torch.randn(G, 192, device="cuda", generator=g), dim=1)
gp_l2t = torch.zeros(192, 192, device="cuda"); gp_l2t[:, :G] = gp.t()
act = ops.new_act(B * N, 4, "cuda")
for impl in ("fp32", "tc"):
    print(json.dumps({"gw": impl, "ms": time_ms(lambda: ops.gw_project(ec, gp_l2t, G, cosine_act=act, kb0=1, impl=impl))}), flush=True)
n, D, K = 1_000_000, 192, 150
cent = torch.randn(K, D, device="cuda", generator=g)
X = cent[torch.randint(0, K, (n,), device="cuda", generator=g)] + 0.35 * torch.randn(n, D, device="cuda", generator=g)
xt = X.t().contiguous()
ct = torch.zeros(D, 152, device="cuda"); ct[:, :K] = (cent + 0.1 * torch.randn(K, D, device="cuda", generator=g)).t()
a = ops.kmeans_assign(xt, ct, K, impl="fp32"); b = ops.kmeans_assign(xt, ct, K, impl="tc")
print(json.dumps({"kmeans_labels_equal": bool(torch.equal(a, b))}))
for impl in ("fp32", "tc"):
    print(json.dumps({"kmeans": impl, "n": n, "ms": time_ms(lambda: ops.kmeans_assign(xt, ct, K, impl=impl))}), flush=True)
c = ops.kmeans_assign_rows(X, ct, K)
print(json.dumps({"kmeans": "tc rows", "n": n, "equal": bool(torch.equal(a, c)), "ms": time_ms(lambda: ops.kmeans_assign_rows(X, ct, K))}), flush=True)
# how many rows does the tensor-core path hand to the pinned re-check?
from gfs3d._lib import lib
def recheck_count_gw():
    nbytes = int(lib().gfs_rowsel_tc_workspace_bytes(B * N, 192))
    ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    assign = torch.empty(B, N, dtype=torch.int32, device="cuda")
    rc = lib().gfs_gw_project_tc(ec.data_ptr(), ec.stride(0), B, 192, N, gp_l2t.data_ptr(), G, 192, act.data_ptr(), act.shape[1], 1, None,
                                 assign.data_ptr(), ws.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    off_cnt = 3 * 2 * 192 * 128 + 256
    return rc, int(ws[off_cnt:off_cnt + 4].view(torch.int32)[0])
print(json.dumps({"gw_recheck_rows": recheck_count_gw(), "rows": B * N}))

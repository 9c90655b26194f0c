"""Where does the end-to-end leg of bench.py lose time?  Variants of the timed loop, device ms/step and host ms/iteration."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gfs-3dseg_gws_b200"))
import bench
from gfs3d.graph import GraphedEval
from gfs3d.synthetic import synthetic_blocks

dev = torch.device("cuda", 0); torch.cuda.set_device(0)
m, gp = bench.build_model(dev)
gened, bc, nc = bench.head_inputs(dev)
B, N = 32, bench.NPTS
host = [synthetic_blocks(B, N, seed=1234 + 100 * i).pin_memory() for i in range(4)]
xs = [h.to(dev) for h in host]
def step(x):
    with torch.no_grad():
        return m(x=x, y=None, eval_model=True, gened_proto=gened, base_class_coding=bc, novel_class_coding=nc)[0]
for i in range(3): step(xs[i])
torch.cuda.synchronize()
g = GraphedEval(step, xs[0])
for i in range(3): g(xs[i])
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
labels_host = torch.empty(B, N, dtype=torch.int32).pin_memory()
K = 30
def run(name, body, do_flush=True):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(K):
        if do_flush: flush.zero_()
        ev[i][0].record(); body(i); ev[i][1].record()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(json.dumps({"variant": name, "dev_ms": sum(a.elapsed_time(b) for a, b in ev) / K, "host_enqueue_ms": 1e3 * (t1 - t0) / K, "wall_ms": 1e3 * (t2 - t0) / K}), flush=True)
run("graph only, resident input", lambda i: g.graph.replay())
run("graph + static_in copy", lambda i: g(xs[i % 4]))
run("graph + argmax + D2H", lambda i: labels_host.copy_(g(xs[i % 4]).argmax(1).to(torch.int32), non_blocking=True))
run("serial H2D + graph + argmax + D2H", lambda i: labels_host.copy_(g(host[i % 4]).argmax(1).to(torch.int32), non_blocking=True))
run("serial, no flush", lambda i: labels_host.copy_(g(host[i % 4]).argmax(1).to(torch.int32), non_blocking=True), do_flush=False)
run("graph only, no flush", lambda i: g.graph.replay(), do_flush=False)

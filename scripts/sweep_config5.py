#!/usr/bin/env python
"""BASELINE.json configs[4]: scaling sweep N in {2048, 4096, 8192} x k in {20, 40}, full GFS eval forward, one GPU.
Prints one JSON line per point (blocks/s with inputs resident, CUDA-event timed, L2 flushed between steps)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gfs-3dseg_gws_b200"))
import bench  # noqa: E402
from gfs3d import ops  # noqa: E402
from gfs3d.synthetic import synthetic_blocks  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gened, bc, nc = bench.head_inputs(dev)
cpu = "--cpu" in sys.argv
for N in (2048, 4096, 8192):
    for k in (20, 40):
        B = 64 if N <= 4096 else 32
        bench.KNN = k
        m, gp = bench.build_model(dev)
        x = synthetic_blocks(B, N, seed=5).to(dev)
        def step():
            with torch.no_grad():
                return m(x=x, y=None, eval_model=True, gened_proto=gened, base_class_coding=bc, novel_class_coding=nc)[0]
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ev = []
        for _ in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record()
            ev.append((e0, e1))
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
        line = {"N": N, "k": k, "batch": B, "ms_per_step": ms, "blocks_per_s": B / ms * 1e3, "points_per_s": B * N / ms * 1e3,
                "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
        if cpu and N <= 4096:
            from oracle import gfs_oracle as O
            sd = {kk: v.detach().float().cpu() for kk, v in m.state_dict().items()}
            xc = x[:2].cpu()
            g2, b2, n2 = bench.head_inputs("cpu")
            t0 = time.perf_counter()
            with torch.no_grad():
                O.forward_eval(sd, gp.cpu(), xc, g2, b2, n2, bench.BASE_NUM, 1.2, k=k)
            line["cpu_port_blocks_per_s"] = 2 / (time.perf_counter() - t0)
            line["cpu_cores"] = os.cpu_count()
        print(json.dumps(line), flush=True)

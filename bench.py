#!/usr/bin/env python
"""bench.py -- DGCNN+GW inference throughput (blocks/s) on synthetic S3DIS-shaped blocks, BASELINE.json configs[1]:
full GFS model eval forward (DGCNN + attention + GW head, 13 classes, 150 geometric words), batch = 32 blocks of
2048 points x 9 channels per GPU, k = 20, random-init weights.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl gfs3d|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one eval forward of the model over one batch of 32 blocks.  Blocks are independent: every
rank processes its own 32 blocks, no data-path collective (weak scaling).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "gfs-3dseg_gws_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "dgcnn_gw_inference_blocks_per_sec"
UNIT = "blocks/s"
CLASSES, BASE_NUM, G, NPTS, KNN = 13, 7, 150, 2048, 20


def model_args():
    return SimpleNamespace(edgeconv_widths=[[64, 64]] * 3, dgcnn_mlp_widths=[512, 256], pc_in_dim=9, dgcnn_k=KNN,
                           base_widths=[128, 64], output_dim=64, eval_weight=1.2)


def head_inputs(device):
    g = torch.Generator().manual_seed(11)
    gened = torch.nn.functional.normalize(torch.randn(CLASSES, 128, generator=g), dim=1)
    coding = (torch.rand(CLASSES, G, generator=g) < 0.3).float()
    return gened.to(device), coding[:BASE_NUM].to(device), coding[BASE_NUM:].to(device)


def build_model(device):
    from gfs3d.synthetic import randomize_bn_
    from model.capl import mpti_net_Point_GeoAsWeight_v2
    import contextlib
    torch.manual_seed(321)
    gp = torch.randn(G, 192, generator=torch.Generator().manual_seed(7))
    with contextlib.redirect_stdout(sys.stderr):      # the constructor prints (as the reference's does); stdout carries ONE JSON line
        m = mpti_net_Point_GeoAsWeight_v2(classes=CLASSES, criterion=torch.nn.CrossEntropyLoss(ignore_index=255), args=model_args(),
                                          base_num=BASE_NUM, gp=gp.to(device), energy=0.9)
    randomize_bn_(m, seed=6)
    return m.to(device).eval(), gp


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_blocks_per_sec(state_dict, gp, sample_blocks, iters, threads):
    """the oracle port (same ATen ops as the reference's CPU path, oracle/gfs_oracle.py) on the host cores"""
    from gfs3d.synthetic import synthetic_blocks
    from oracle import gfs_oracle as O
    torch.set_num_threads(threads)
    sd = {k: v.detach().float().cpu() for k, v in state_dict.items()}
    gened, bc, nc = head_inputs("cpu")
    x = synthetic_blocks(sample_blocks, NPTS, seed=999)
    with torch.no_grad():
        O.forward_eval(sd, gp.cpu(), x[:1], gened, bc, nc, BASE_NUM, 1.2, k=KNN)     # warm-up
        t0 = time.perf_counter()
        for _ in range(iters):
            O.forward_eval(sd, gp.cpu(), x, gened, bc, nc, BASE_NUM, 1.2, k=KNN)
        dt = time.perf_counter() - t0
    return sample_blocks * iters / dt


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path (Python/ATen; restated in oracle/ because
    /root/reference cannot travel to the GPU box), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    m, gp = build_model("cpu")
    sample = 4
    t0 = time.perf_counter()
    v = cpu_port_blocks_per_sec(m.state_dict(), gp, sample, max(1, a.steps), threads)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * sample / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"GFS eval forward, S3DIS-shaped blocks N={NPTS} C=9 k={KNN}, {CLASSES} classes, {G} GWs",
                       "blocks_per_step": sample, "note": "CPU path, bounded sample of the batch=32 workload"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": f"{sample} blocks x {max(1, a.steps)} steps"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="gfs3d", choices=["gfs3d", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="blocks per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--e2e-serial", action="store_true",
                    help="end-to-end leg without prefetch: H2D, forward and D2H of a step strictly one after the other")
    a = ap.parse_args()
    if a.impl == "reference":
        return run_reference(a)
    a.warmup = max(a.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from gfs3d import ops
    from gfs3d.synthetic import synthetic_blocks
    m, gp = build_model(dev)
    gened, bc, nc = head_inputs(dev)
    B = a.batch
    NROT = 4     # distinct input batches, rotated
    host = [synthetic_blocks(B, NPTS, seed=1234 + 1000 * rank + 100 * i).pin_memory() for i in range(NROT)]
    xs = [h.to(dev) for h in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    labels_host = torch.empty(B, NPTS, dtype=torch.int32).pin_memory()

    def step(x):
        with torch.no_grad():
            logits, _, _ = m(x=x, y=None, eval_model=True, gened_proto=gened, base_class_coding=bc, novel_class_coding=nc)
        return logits

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(a.warmup):
        step(xs[i % NROT])
    barrier()
    graphed = None
    if not a.no_graph:
        from gfs3d.graph import GraphedEval
        graphed = GraphedEval(step, xs[0])
        for i in range(a.warmup):
            graphed(xs[i % NROT])
        barrier()
    run = graphed if graphed is not None else step

    # ---- timed region 1: inputs resident in HBM; per-step CUDA events, L2 flushed between steps ----
    clocks = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    l0 = ops.LAUNCHES
    barrier()
    w0 = time.perf_counter()
    for i in range(a.steps):
        flush.zero_()
        ev[i][0].record()
        run(xs[i % NROT])
        ev[i][1].record()
    barrier()
    wall = time.perf_counter() - w0
    launches = (ops.LAUNCHES - l0) if graphed is None else graphed.kernels_per_replay * a.steps
    dev_ms = sum(s.elapsed_time(e) for s, e in ev)

    # ---- timed region 2: end to end through the public module API with HOST buffers (H2D + forward + D2H labels) ----
    # Default: the next batch's pinned host -> device copy is issued (on a copy stream, into a second device buffer) at the
    # START of a step's timed interval, so it runs under that step's kernels; every interval contains exactly one H2D, one
    # forward and one D2H of the labels.  --e2e-serial keeps the three strictly one after the other.
    xdev = [torch.empty(B, 9, NPTS, device=dev) for _ in range(2)]
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]     # H2D into xdev[j] finished
    freed = [torch.cuda.Event(), torch.cuda.Event()]     # the forward that read xdev[j] finished
    barrier()
    if not a.e2e_serial:
        with torch.cuda.stream(copy_stream):             # prologue: the first batch (its copy is the one step K-1 would prefetch)
            xdev[0].copy_(host[0], non_blocking=True)
            ready[0].record(copy_stream)
        freed[1].record(main_stream)
    for i in range(a.steps):
        flush.zero_()
        ev2[i][0].record()
        if a.e2e_serial:
            if graphed is not None:
                lg = graphed(host[i % NROT])                 # pinned host -> static device input, then one graph launch
            else:
                xdev[0].copy_(host[i % NROT], non_blocking=True)
                lg = step(xdev[0])
        else:
            cur, nxt = i & 1, (i + 1) & 1
            copy_stream.wait_event(ev2[i][0])                # the prefetch belongs to this interval
            copy_stream.wait_event(freed[nxt])
            with torch.cuda.stream(copy_stream):
                xdev[nxt].copy_(host[(i + 1) % NROT], non_blocking=True)
                ready[nxt].record(copy_stream)
            main_stream.wait_event(ready[cur])
            lg = graphed(xdev[cur]) if graphed is not None else step(xdev[cur])
            freed[cur].record(main_stream)
        labels_host.copy_(lg.argmax(1).to(torch.int32), non_blocking=True)
        if not a.e2e_serial:
            main_stream.wait_event(ready[(i + 1) & 1])       # the interval ends after ITS host -> device copy as well
        ev2[i][1].record()
    barrier()
    e2e_ms = sum(s.elapsed_time(e) for s, e in ev2)
    clk = clocks.stop() if clocks else None

    # ---- instrumented replay of the same steps: per-entry-point CUDA events for the roofline ----
    ops.PROFILE = {}
    for i in range(a.steps):
        flush.zero_()
        step(xs[i % NROT])
    torch.cuda.synchronize()
    prof = {k: sum(s.elapsed_time(e) for s, e in v) / a.steps for k, v in ops.PROFILE.items()}   # ms per step per entry point
    calls = {k: len(v) // a.steps for k, v in ops.PROFILE.items()}
    ops.PROFILE = None

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"

    # dominant kernel: the fused kNN (3 launches per step; the two C=64 layers dominate)
    knn_ms = prof.get("gfs_knn_tc_f32", 0.0) + prof.get("gfs_knn_f32", 0.0)
    knn_bytes = sum(B * NPTS * (c * 4 + KNN * 4) for c in (9, 64, 64))                  # read x once + write idx, per step
    knn_flops = sum(2.0 * c * NPTS * NPTS * B for c in (9, 64, 64))
    ec_ms = prof.get("gfs_edgeconv_fwd", 0.0)
    ec_bytes = 3 * B * NPTS * (128 * 4 + KNN * 4 + 64 * 4)                              # read P',Q',idx + write y
    ec123_ms = knn_ms + ec_ms + prof.get("gfs_pointwise_f32", 0.0)
    ec123_bytes = 1316 * NPTS * B                                                        # compulsory bytes
    gbs = lambda byt, ms: (byt / 1e9) / (ms / 1e3) if ms > 0 else None
    traffic = None          # dram__bytes_read+write of the kNN kernels of a step (3 x prep, filter, finish), from the committed ncu captures
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["knn_bytes_per_step_b32"] * B / 32.0
    except Exception:
        pass
    roof = {"kernel": "kNN graph: knn_prep + knn_tc (tcgen05 filter) + knn_finish (gfs_knn_tc_f32, 3 calls/step)", "bound": "hbm", "achieved": gbs(knn_bytes, knn_ms), "peak": hbm_peak,
            "unit": "GB/s", "frac": (gbs(knn_bytes, knn_ms) or 0) / hbm_peak, "traffic": traffic, "peak_source": peak_src,
            "ms_per_step": knn_ms, "share_of_step": knn_ms / (sum(prof.values()) or 1),
            "binding_roof": "per-row top-k selection out of TMEM (dependent-issue latency), not HBM or the tensor pipe",
            "distance_tflops_algorithmic": knn_flops / 1e12 / (knn_ms / 1e3) if knn_ms else None,
            "timing": "CUDA events around every C-ABI call on the launch stream, instrumented replay of the same K steps"}
    extra = {
        "edgeconv_given_graph": {"bound": "hbm", "achieved": gbs(ec_bytes, ec_ms), "peak": hbm_peak, "unit": "GB/s",
                                 "frac": (gbs(ec_bytes, ec_ms) or 0) / hbm_peak, "ms_per_step": ec_ms},
        "edgeconv123_fused_knn": {"bound": "hbm", "achieved": gbs(ec123_bytes, ec123_ms), "peak": hbm_peak, "unit": "GB/s",
                                  "frac": (gbs(ec123_bytes, ec123_ms) or 0) / hbm_peak, "ms_per_step": ec123_ms,
                                  "note": "compulsory bytes 1316*N per block; ceiling with brute-force fp32 kNN is ~2.5 % (SURVEY 8d)"},
        "entry_point_ms_per_step": prof, "entry_point_calls_per_step": calls,
    }

    cpu = None
    if not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample, iters = 4, 3
        v = cpu_port_blocks_per_sec(m.state_dict(), gp, sample, iters, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": f"{sample} blocks x {iters} iterations of the same workload"}

    total_blocks = B * a.steps * world
    h2d = B * 9 * NPTS * 4
    d2h = B * NPTS * 4
    line = {"metric": METRIC, "value": total_blocks / (dev_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+bf16",
            "data": "synthetic",
            "config": {"workload": f"BASELINE.json configs[1]: full GFS eval forward, batch={B} S3DIS-shaped blocks/GPU, N={NPTS}, C=9, k={KNN}, "
                                   f"{CLASSES} classes, {G} GWs, random-init weights", "blocks_per_gpu_per_step": B,
                       "l2": "flushed between steps (256 MiB write outside the per-step event pair); 4 rotating input batches",
                       "parallelism": f"block-sharded x{world}, no data-path collective",
                       "launch": "eager" if graphed is None else "CUDA graph replay of the eager step (gfs3d/graph.py)",
                       "attention": "hand-written tcgen05 flash kernel (gfs_attention_fwd)"},
            "clocks": clk, "e2e": {"value": total_blocks / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                   "ms_per_step": e2e_ms / a.steps,
                                   "mode": "serial: H2D, forward, D2H one after the other" if a.e2e_serial else
                                           "prefetch: each step's timed interval holds one pinned H2D (the next batch, on a copy "
                                           "stream, into a second device buffer), one forward and one D2H of the labels"},
            "gpu_launches": launches, "roofline": roof, "roofline_detail": extra, "cpu_baseline": cpu, "wall_s_timed_region": wall}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

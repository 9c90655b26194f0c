#!/usr/bin/env python
"""Build profiles/<round>_summary.md from the files scripts/gpu_profile.sh leaves in gpurun_out/ (run here, no GPU)."""
import collections
import csv
import json
import os
import subprocess
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
out = [f"# Round {R[1:]} profile summary (B200, `scripts/gpu_profile.sh {R}`)\n"]

# ---- bench lines
for name in (f"{R}_bench_n1.json", f"{R}_bench_reference.json"):
    fn = os.path.join(G, name)
    if os.path.exists(fn):
        line = [l for l in open(fn).read().strip().splitlines() if l.startswith("{")][-1]
        open(os.path.join(P, name), "w").write(line + "\n")
        d = json.loads(line)
        out.append(f"* `{name}`: **{d['value']:.1f} {d['unit']}** ({d.get('ms_per_step', 0):.3f} ms/step), e2e {d['e2e']['value']:.1f}"
                   + (f", clocks {d['clocks']}" if d.get("clocks") else ""))
        if d.get("cpu_baseline"):
            out.append(f"  * cpu_baseline: {d['cpu_baseline']}")

# ---- launch list: one step's kernels and their share
fn = os.path.join(G, f"{R}_launches.csv")
if os.path.exists(fn):
    rows = [r for r in csv.reader(open(fn)) if len(r) > 10]
    hdr, data = rows[0], rows[1:]
    iN, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = [(r[iN], float(r[iV]) / 1000.0) for r in data]
    # a step starts with the first kNN preparation launch that follows the head's last kernel (the second cos_logits launch)
    # and runs up to the next such launch; a kNN call is one chain of launches or two (gfs_knn_tc_chains)
    starts = [i for i, (n, _) in enumerate(seq)
              if ("sqnorm_kernel" in n or "knn_prep_kernel" in n) and i > 0 and not any(t in seq[i - 1][0] for t in ("knn_", "edgeconv", "edge_pq", "sqnorm"))]
    if len(starts) >= 2:
        step = seq[starts[0]:starts[1]]
        tot = sum(t for _, t in step)
        agg = collections.OrderedDict()
        for n, t in step:
            key = n.split("(")[0].replace("void ", "")
            key = key if key.startswith("gfs::") else "torch/ATen: " + key[:60]
            a = agg.setdefault(key, [0, 0.0])
            a[0] += 1
            a[1] += t
        out.append(f"\n## Launch list of one timed step (ncu gpu__time_duration, cold cache, serialised): {len(step)} launches, {tot:.0f} us\n")
        out.append("| kernel | launches | us | share |\n|---|---:|---:|---:|")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if t / tot >= 0.002:
                out.append(f"| `{k}` | {c} | {t:.1f} | {100 * t / tot:.1f} % |")
        ours = sum(t for k, (c, t) in agg.items() if k.startswith("gfs::"))
        out.append(f"\nhand-written kernels: {100 * ours / tot:.1f} % of the step's device time")
    with open(os.path.join(P, f"{R}_launches.csv"), "w") as f:
        f.write(open(fn).read())

# ---- per-kernel ncu --set full metrics
WANT = [("gpu__time_duration.sum", "duration"), ("launch__registers_per_thread", "regs/thread"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %")]
out.append("\n## ncu --set full, one capture per hand-written kernel\n")
for fn in sorted(os.listdir(G)):
    if not (fn.startswith(f"{R}_raw_") and fn.endswith(".csv")):
        continue
    rows = list(csv.reader(open(os.path.join(G, fn))))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        vals = []
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                vals.append(f"{label} {r[i]} {units[i]}".strip())
        out.append(f"* `{name}`: " + "; ".join(vals))

# ---- roofline table: achieved bandwidth from the ALGORITHMIC bytes of the bench shape (B = 32, N = 2048, k = 20) and from the
# measured dram traffic, against the measured HBM peak; tensor-pipe utilisation for the GEMM-shaped kernels
peaks = {}
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peaks = json.load(open(pk))
HBM = float(peaks.get("hbm_gbs", 6547.2))
Bn = 32 * 2048
ALG = {   # kernel -> algorithmic bytes per launch, by the launch's position in the capture
    "knn_prep_kernel": [Bn * (4 * 9), Bn * (4 * 64), Bn * (4 * 64)],
    "knn_tc_kernel": [Bn * (4 * 9 + 80), Bn * (4 * 64 + 80), Bn * (4 * 64 + 80)],
    "knn_finish_kernel": [Bn * 80, Bn * 80, Bn * 80],
    "edgeconv_kernel": [Bn * (128 + 256 + 80 + 256 + 128)],
    "edge_pq_kernel": [Bn * (4 * 64 + 128 + 256)],
    "rowsel_tc_kernel": [Bn * (192 * 4 + 160 * 2 + 4)],
    "cos_logits_kernel": [Bn * (128 * 4 + 13 * 4)],
    "softmax_pool_kernel": [Bn * (128 * 4 + 13 * 4)],
    "attention_kernel": [Bn * (192 * 2 + 64 * 4 + 128)],
}
out.append(f"\n## Roofline per kernel (bench shape; HBM peak {HBM:.0f} GB/s measured)\n")
out.append("| kernel (launch) | duration us | algorithmic MB | algorithmic GB/s | % of HBM peak | measured dram MB | dram GB/s | tensor pipe % | issue active % |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|")
for fn in sorted(os.listdir(G)):
    if not (fn.startswith(f"{R}_raw_") and fn.endswith(".csv")):
        continue
    rows = list(csv.reader(open(os.path.join(G, fn))))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]

    def val(r, key):
        if key not in hdr:
            return None
        i = hdr.index(key)
        v = float(r[i].replace(",", ""))
        u = units[i]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u, 1.0)
    for li, r in enumerate(rows[2:]):
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        base = name.split("<")[0]
        dur = val(r, "gpu__time_duration.sum")
        dram = (val(r, "dram__bytes_read.sum") or 0) + (val(r, "dram__bytes_write.sum") or 0)
        alg = ALG.get(base)
        a = alg[li] if alg and li < len(alg) else None
        tp = val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        ia = val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
        out.append(f"| `{name}` ({li}) | {dur:.1f} | " + (f"{a / 1e6:.1f} | {a / dur / 1e3:.0f} | {100 * a / dur / 1e3 / HBM:.1f} % | " if a else "- | - | - | ")
                   + f"{dram / 1e6:.1f} | {dram / dur / 1e3:.0f} | {tp:.1f} | {ia:.1f} |")

# ---- dram traffic of the dominant entry point (the kNN graph: prep + filter + finish, three launches each per step)
traffic = {}
for kname in ("knn_prep_kernel", "knn_tc_kernel", "knn_finish_kernel"):
    fn = os.path.join(G, f"{R}_raw_{kname}.csv")
    if not os.path.exists(fn):
        continue
    rows = list(csv.reader(open(fn)))
    hdr, units = rows[0], rows[1]

    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    traffic[kname] = [[to_bytes(r[ir], units[ir]), to_bytes(r[iw], units[iw])] for r in rows[2:] if len(r) > max(ir, iw)]
if len(traffic) == 3 and all(len(v) in (3, 6) for v in traffic.values()):
    total = sum(a + b for v in traffic.values() for a, b in v)
    json.dump({"source": f"profiles/{R}_summary.md (ncu --set full): dram__bytes_read.sum + dram__bytes_write.sum of the three launches "
                         "(layers 1-3, one or two chains per layer) of knn_prep_kernel, knn_tc_kernel and knn_finish_kernel of one step at batch 32",
               "per_launch_bytes_read_write": traffic, "knn_bytes_per_step_b32": total},
              open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
    out.append(f"\nkNN graph dram traffic per step (batch 32): {total / 1e6:.1f} MB "
               f"(algorithmic: 51.6 MB = read the three layer inputs once + write the index lists)")

# ---- SASS evidence
so = os.path.join(ROOT, "gfs-3dseg_gws_b200", "gfs3d", "libgfs3d.so")
if os.path.exists(so):
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            counts[cur] = collections.Counter()
        elif cur:
            for mn in ("UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "LDGSTS", "SYNCS", "FFMA", "HMMA", "SHFL", "STS", "LDS"):
                if f" {mn}" in line or f"{mn}." in line:
                    counts[cur][mn] += 1
    out.append("\n## SASS mnemonics per kernel (cuobjdump -sass libgfs3d.so; UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = TMA bulk copy, LDGSTS = cp.async)\n")
    out.append("| kernel | UTCHMMA | LDTM | UBLKCP | LDGSTS | SYNCS (mbarrier) | FFMA | HMMA (legacy) |\n|---|---:|---:|---:|---:|---:|---:|---:|")
    for k, c in counts.items():
        short = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
        out.append(f"| `{short}` | {c['UTCHMMA']} | {c['LDTM']} | {c['UBLKCP']} | {c['LDGSTS']} | {c['SYNCS']} | {c['FFMA']} | {c['HMMA']} |")
    for fn in sorted(os.listdir(G)):                      # keep the raw ncu pages next to the summary
        if fn.startswith(f"{R}_raw_") and fn.endswith(".csv"):
            with open(os.path.join(P, fn), "w") as f:
                f.write(open(os.path.join(G, fn)).read())
    open(os.path.join(P, f"{R}_sass_listing.txt"), "w").write(
        "\n".join(l for l in sass.splitlines() if any(m in l for m in ("Function :", "UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "LDGSTS"))) + "\n")

open(os.path.join(P, f"{R}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:6000])

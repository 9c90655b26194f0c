#!/usr/bin/env python
"""Hot spots of an `ncu --page source --csv` export: stall samples per SASS instruction, grouped by how often the instruction
ran (= which warp role it belongs to), with the dominant stall reasons.  usage: src_hotspots.py file.csv [section index]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lo, hi = starts[k], (starts[k + 1] if k + 1 < len(starts) else len(rows))
print(rows[lo][1][:120])
h = rows[lo + 1]
iS, iA, iI = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
st = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
data = []
for r in rows[lo + 2:hi]:
    try:
        data.append((int(r[iS]), int(r[iI] or 0), r[iA], {c: int(r[h.index(c)] or 0) for c in st}))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
seg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for s, e, src, stl in data:
    g = seg[e]
    g[0] += 1
    g[1] += s
    g[2].update(stl)
print(f"total samples {tot}")
for e, (n, s, c) in sorted(seg.items(), key=lambda kv: -kv[1][1])[:10]:
    print(f"  executed {e:>9} x: {n:5d} instructions, {s:6d} samples ({100 * s / tot:4.1f} %)  {c.most_common(3)}")
for s, e, src, stl in sorted(data, key=lambda d: -d[0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    print(f"{s:6d} {100 * s / tot:5.1f}% x{e:<9} {src[:64]:64s} {sorted(stl.items(), key=lambda kv: -kv[1])[:2]}")

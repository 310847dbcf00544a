#!/usr/bin/env python
"""Per-launch DRAM traffic of a kernel sequence from an ncu CSV (gpu__time_duration.sum, dram__bytes_read/write.sum):
    python tools/dram_sequence.py gpurun_out/dram_backward_warm_r2i.csv > profiles/dram_backward_warm_r2i.md
Run ncu with --cache-control none to keep the L2 contents of the preceding launch (the kernels are still serialised)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv, mn, idc, mu = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "ID", "Metric Unit"))
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("pcg::", "")
    d = per.setdefault(r[idc], {"name": name})
    v = float(r[mv].replace(",", ""))
    if r[mn].startswith("dram__bytes"):
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[r[mu]]
    elif r[mn] == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1e-3)
    d[r[mn]] = v
seq = list(per.values())
print(f"# DRAM traffic per launch, in launch order: {sys.argv[1]}\n")
print("| kernel | us | DRAM read MB | DRAM write MB |\n|---|---|---|---|")
agg = collections.OrderedDict()
for d in seq:
    t, rd, wr = (d.get(k, 0.0) for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"))
    print(f"| `{d['name'][:60]}` | {t:.1f} | {rd:.1f} | {wr:.1f} |")
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] += rd; a[3] += wr
print("\n## averages\n\n| kernel | launches | us | DRAM read MB | DRAM write MB |\n|---|---|---|---|---|")
for k, (n, t, rd, wr) in agg.items():
    print(f"| `{k[:60]}` | {n} | {t / n:.1f} | {rd / n:.1f} | {wr / n:.1f} |")

"""Instruction-count table of the shipped library: which kernels use tcgen05 (UTCHMMA = tcgen05.mma kind::f16, LDTM / STTM
= tcgen05.ld / st on TMEM), TMA (UTMALDG / UTMASTG / UTMAPF), legacy warp-level tensor-core HMMA (mma.sync), clusters.
    python tools/sass_table.py > profiles/sass_counts_r2.md      (needs cuobjdump; no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "promptable-counterfactual-gan_b200", "csrc", "libpcg.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "HMMA", "SYNCS", "UCGABAR", "FFMA", "DFMA"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("pcg::", "")
        cur = counts.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        for k in KEYS:
            if op.startswith(k):
                cur[k] += 1
print("# SASS instruction counts of libpcg.so (sm_100a), `cuobjdump -sass`, per kernel\n")
print("UTCHMMA = tcgen05.mma (kind::f16), LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG = TMA load/store, "
      "UTCBAR = tcgen05.commit, HMMA = mma.sync, SYNCS = mbarrier ops, UCGABAR = cluster barrier.  Only kernels with at least one of "
      "these are listed.\n")
print("| kernel | " + " | ".join(KEYS) + " | total |")
print("|---|" + "---|" * (len(KEYS) + 1))
for name, c in counts.items():
    if not any(c[k] for k in KEYS[:11]):
        continue
    print(f"| `{name[:70]}` | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + f" | {c['total']} |")
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("\nLibrary totals: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]) + f"; {len(counts)} kernels.")

#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_rN.csv  > profiles/launches_rN.md
    python tools/summarize_ncu.py kernel   gpurun_out/prof_X.ncu-rep   > profiles/prof_X.md
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg, tot, n = collections.OrderedDict(), 0.0, 0
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("pcg::", "")
        t = float(r[mv].replace(",", "")) / 1000.0
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
        n += 1
    print(f"# ncu launch list: {path}\n")
    print(f"`--metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES). "
          f"{n} launches, {tot / 1000:.2f} ms total.\n")
    print("| share | total us | launches | avg us | kernel |\n|---|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {100 * t / tot:.1f}% | {t:.1f} | {c} | {t / c:.1f} | `{k[:100]}` |")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full: {path}\n")
    for r in rows[2:]:
        print(f"## `{r[hdr.index('Kernel Name')][:120]}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"| {w} | {r[i]} | {units[i]} |")
        print()
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hi = [i for i, r in enumerate(rows) if "# Samples" in r]
    if hi:
        hdr = rows[hi[0]]
        S, SRC = hdr.index("# Samples"), hdr.index("Source")
        stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        data = [r for r in rows[hi[0] + 1:] if len(r) > S and r[S].isdigit()]
        tot = sum(int(r[S]) for r in data) or 1
        print("## hottest SASS lines (warp-stall samples)\n\n| samples | share | instruction | top stalls |\n|---|---|---|---|")
        for r in sorted(data, key=lambda r: -int(r[S]))[:12]:
            st = sorted(((hdr[i][6:], int(r[i])) for i in stall if int(r[i]) > 0), key=lambda kv: -kv[1])[:3]
            print(f"| {r[S]} | {100 * int(r[S]) / tot:.1f}% | `{r[SRC].strip()[:70]}` | {st} |")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])

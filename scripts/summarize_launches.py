"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per (kernel, grid, block) launches, total and mean time, share.
    python scripts/summarize_launches.py gpurun_out/launches.csv profiles/r01_launches_XXX.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
iK, iG, iB, iM, iV, iU = (h.index(k) for k in ("Kernel Name", "Grid Size", "Block Size", "Metric Name", "Metric Value", "Metric Unit"))
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[iM] != "gpu__time_duration.sum":
        continue
    v = float(r[iV].replace(",", ""))
    us = v / 1e3 if r[iU] in ("ns", "nsecond") else (v * 1e3 if r[iU] in ("ms", "msecond") else v)
    name = re.sub(r"\(.*$", "", r[iK]).replace("void ", "")
    key = (name, r[iG], r[iB])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
w = csv.writer(open(sys.argv[2], "w"))
w.writerow(["kernel", "grid", "block", "launches", "total_us", "avg_us", "share"])
for (k, g, b), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    w.writerow([k, g, b, n, f"{t:.1f}", f"{t / n:.2f}", f"{t / tot:.4f}"])
print(f"{sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time")

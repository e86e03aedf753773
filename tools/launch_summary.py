"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and average
time, share.  usage: launch_summary.py launches.csv [skip_first_n_launches]"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, r.get("Grid Size", ""), r.get("Block Size", ""), v))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else len(rows) // 2      # default: the second of two iterations
rows = rows[skip:]
agg = OrderedDict()
for name, g, b, v in rows:
    k = (name, g, b)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot:.0f} us")
for (name, g, b), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / tot * 100:5.1f} %  {n:4d} x {t / n:8.1f} us  {name[:90]} {g} x {b}")

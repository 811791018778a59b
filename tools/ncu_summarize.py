"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launches per kernel."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(unit, 1e-3)
    rows.append((r["Kernel Name"], v))
tot = sum(v for _, v in rows)
agg = defaultdict(lambda: [0, 0.0])
for k, v in rows:
    k = re.sub(r"\(.*", "", k)
    agg[k][0] += 1
    agg[k][1] += v
print(f"{len(rows)} launches, {tot/1e3:.2f} ms total (serialised, cold cache: compare shares)")
print(f"{'kernel':70s} {'launches':>8s} {'ms':>9s} {'share':>7s} {'avg us':>9s}")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:8d} {v/1e3:9.2f} {100*v/tot:6.1f}% {v/n:9.1f}")

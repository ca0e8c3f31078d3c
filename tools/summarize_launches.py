"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
    rows.append((r["Kernel Name"], v * scale))
agg = defaultdict(lambda: [0, 0.0])
for name, us in rows:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"<lambda.*", "<lambda>", short)
    if len(short) > 110:
        short = short[:110]
    agg[short][0] += 1
    agg[short][1] += us
total = sum(v[1] for v in agg.values())
print("total launches %d, total kernel time %.1f ms (cold-cache, serialised: compare shares)" % (len(rows), total / 1e3))
print("%-112s %7s %10s %7s %9s" % ("kernel", "calls", "ms", "share", "us/call"))
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-112s %7d %10.3f %6.1f%% %9.1f" % (name, n, us / 1e3, 100 * us / total, us / n))

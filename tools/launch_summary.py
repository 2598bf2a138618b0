"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py file.csv"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].split("(")[0][-64:]
    v, u = float(r[ix["Metric Value"]]), r[ix["Metric Unit"]]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print("%-66s launches %4d %10.3f ms %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
print("total %.3f ms over %d launches" % (tot, sum(v[0] for v in agg.values())))

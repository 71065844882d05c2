"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv
import sys
from collections import defaultdict

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    k = r["Kernel Name"].split("(")[0].replace("void ", "")[:70]
    agg[k][0] += 1
    agg[k][1] += float(r["Metric Value"].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {len(rows)} launches, {tot/1e6:.3f} ms total (ncu-serialised, cold cache)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} n={v[0]:3d} total_ms={v[1]/1e6:10.3f} share={v[1]/tot:6.3f}")

"""Dynamic SASS opcode mix and stall-sample hot spots from `ncu --page source --csv` output."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = defaultdict(float)
samples = defaultdict(float)
tot = 0.0
body = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    n = float(r[ix["Instructions Executed"]] or 0)
    s = float(r[ix["# Samples"]] or 0)
    toks = src.split()
    op = toks[0]
    if op.startswith("@"):
        op = toks[1]
    op = op.split(".")[0]
    ops[op] += n
    samples[op] += s
    tot += n
    body.append((n, s, src))
print(f"total warp instructions executed: {tot:.4g}")
div = float(sys.argv[2]) if len(sys.argv) > 2 else None
for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:28]:
    extra = f"  per-warp-element {n/div:7.2f}" if div else ""
    print(f"  {op:10s} {n:14.4g}  {100*n/tot:5.1f}%  stall-samples {samples[op]:8.0f}{extra}")
print("-- top stall-sample instructions")
for n, s, src in sorted(body, key=lambda t: -t[1])[:25]:
    print(f"  {s:7.0f}  exec {n:12.4g}  {src[:90]}")

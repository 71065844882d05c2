"""Group the SASS lines of `ncu --page source --csv` output by execution count (= code region)."""
import csv
import sys
from collections import Counter, defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
groups = defaultdict(lambda: [0, 0.0, Counter(), 0.0])
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    n = float(r[ix["Instructions Executed"]] or 0)
    s = float(r[ix["# Samples"]] or 0)
    toks = r[ix["Source"]].strip().split()
    op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
    g = groups[round(n / 1e6, 0)]
    g[0] += 1
    g[1] += n
    g[2][op] += 1
    g[3] += s
    tot += n
alls = sum(g[3] for g in groups.values())
print(f"total warp instructions {tot:.4g}")
for k in sorted(groups, key=lambda k: -groups[k][1])[:12]:
    g = groups[k]
    print(f"exec~{k:6.0f}e6 ninstr={g[0]:4d} total={g[1]:.3g} ({100 * g[1] / tot:4.1f}%) stall {100 * g[3] / alls:4.1f}%  "
          f"{g[2].most_common(7)}")

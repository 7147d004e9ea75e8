#!/usr/bin/env python
"""profiles/launch_summary.py launches.csv -- per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hi]
kn, mv, mu, gs = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Grid Size")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu], 1.0)
    agg.setdefault(r[kn], []).append((v, r[gs]))
tot = sum(x[0] for v in agg.values() for x in v)
print(f"{len(rows) - hi - 1} launches, {tot:.3f} ms device time in total (serialised, cold cache)")
for k, v in sorted(agg.items(), key=lambda kv: -sum(x[0] for x in kv[1])):
    s = sum(x[0] for x in v)
    print(f"{s / tot * 100:6.2f}% {s:10.3f} ms  n={len(v):3d}  {k[:90]}  each(ms)={[round(x[0], 3) for x in v[:10]]}")

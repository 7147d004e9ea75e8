#!/usr/bin/env python
"""profiles/trace_summary.py trace.json -- what a CUPTI timeline (bench.py --e2e-profile) says about overlap: GPU-busy time
(union of kernel intervals), copy time per direction, time with neither, per-kernel totals, idle gaps.  Tooling for profiles/."""
import collections
import gzip
import json
import sys

path = sys.argv[1]
raw = (gzip.open(path, "rt") if path.endswith(".gz") else open(path)).read()
ev = [e for e in json.loads(raw)["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
if not ev:
    sys.exit("no GPU events")
t0 = min(e["ts"] for e in ev)
t1 = max(e["ts"] + e["dur"] for e in ev)


def union(iv):
    iv = sorted(iv)
    tot, cur_s, cur_e = 0.0, None, None
    out = []
    for s, e in iv:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                tot += cur_e - cur_s
                out.append((cur_s, cur_e))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        tot += cur_e - cur_s
        out.append((cur_s, cur_e))
    return tot, out


kern = [(e["ts"], e["ts"] + e["dur"]) for e in ev if e["cat"] == "kernel"]
h2d = [(e["ts"], e["ts"] + e["dur"]) for e in ev if e["cat"] == "gpu_memcpy" and "HtoD" in e["name"]]
d2h = [(e["ts"], e["ts"] + e["dur"]) for e in ev if e["cat"] == "gpu_memcpy" and "DtoH" in e["name"]]
ku, kiv = union(kern)
hu, _ = union(h2d)
du, _ = union(d2h)
au, aiv = union(kern + h2d + d2h)
print(f"span {(t1 - t0) / 1e3:.2f} ms; kernels busy {ku / 1e3:.2f} ms; H2D busy {hu / 1e3:.2f} ms; D2H busy {du / 1e3:.2f} ms; anything busy {au / 1e3:.2f} ms")
hb = sum(e["args"].get("bytes", 0) for e in ev if e["cat"] == "gpu_memcpy" and "HtoD" in e["name"])
db = sum(e["args"].get("bytes", 0) for e in ev if e["cat"] == "gpu_memcpy" and "DtoH" in e["name"])
print(f"H2D {hb / 1e6:.1f} MB at {hb / max(hu, 1) / 1e3:.1f} GB/s while active; D2H {db / 1e6:.1f} MB at {db / max(du, 1) / 1e3:.1f} GB/s while active")
gaps = sorted(((b[0] - a[1]) for a, b in zip(kiv, kiv[1:])), reverse=True)
print("largest kernel-idle gaps (ms):", [round(g / 1e3, 3) for g in gaps[:12]], "sum", round(sum(gaps) / 1e3, 2))
agg = collections.defaultdict(lambda: [0.0, 0])
for e in ev:
    if e["cat"] == "kernel":
        agg[e["name"][:70]][0] += e["dur"]
        agg[e["name"][:70]][1] += 1
for k, (d, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"{d / 1e3:9.3f} ms  n={n:4d}  {k}")
# concurrency: time-weighted number of kernels in flight
pts = sorted([(s, 1) for s, _ in kern] + [(e, -1) for _, e in kern])
lvl, last, hist = 0, pts[0][0], collections.Counter()
for t, d in pts:
    hist[lvl] += t - last
    last = t
    lvl += d
print("kernels in flight (ms):", {k: round(v / 1e3, 2) for k, v in sorted(hist.items())})

#!/usr/bin/env python
"""profiles/launch_metrics.py launches.csv [traffic.json CONFIG] -- table of an ncu multi-metric launch list (one row per launch:
time, warp-instructions, issue-active %, warps per scheduler, DRAM bytes) and per-kernel totals of the LAST step in the file;
optionally writes the DRAM traffic of that step into profiles/r02_traffic.json under "config<CONFIG>" (what bench.py cites as
`roofline.traffic`, with this file as the source)."""
import collections
import csv
import json
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
idx = {h: i for i, h in enumerate(rows[0])}
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[idx["ID"]], {"name": r[idx["Kernel Name"]]})[r[idx["Metric Name"]]] = (r[idx["Metric Value"]], r[idx["Metric Unit"]])


def f(v, m):
    x, u = v.get(m, ("0", ""))
    x = float(x.replace(",", ""))
    return x * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


def short(n):
    n = n.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    return n.split("(")[0][:44]


launches = [(short(v["name"]), f(v, "gpu__time_duration.sum"), f(v, "smsp__inst_executed.sum"), f(v, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
             f(v, "smsp__warps_active.avg.per_cycle_active"), f(v, "dram__bytes_read.sum"), f(v, "dram__bytes_write.sum"), v.get("launch__grid_size", ("", ""))[0],
             v.get("launch__registers_per_thread", ("", ""))[0]) for v in d.values()]
# the last step = everything after the last plan_kernel's preceding lookup (take the launches from the last kmc_hash_lookup_kernel<0> before the last plan_kernel)
last_plan = max(i for i, l in enumerate(launches) if l[0].startswith("plan_kernel"))
start = max((i for i, l in enumerate(launches[:last_plan]) if l[0].startswith("pfkmc::kmc_hash_lookup_kernel") or l[0].startswith("kmc_hash_lookup_kernel")), default=last_plan)
step = launches[start:]
print(f"{len(launches)} launches in the file; last step = launches {start} .. {len(launches) - 1}")
agg = collections.OrderedDict()
for l in step:
    a = agg.setdefault(l[0], [0, 0.0, 0.0, 0.0, 0.0, 0.0, l[7], l[8]])
    a[0] += 1; a[1] += l[1]; a[2] += l[2]; a[3] += l[3] * l[1]; a[4] += l[5]; a[5] += l[6]
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':46s} {'n':>3s} {'ms':>8s} {'share':>6s} {'Mwarp-inst':>10s} {'issue%':>6s} {'dramR MB':>9s} {'dramW MB':>9s}  grid/regs of the first")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:46s} {a[0]:3d} {a[1]:8.3f} {100 * a[1] / tot:5.1f}% {a[2] / 1e6:10.1f} {a[3] / max(a[1], 1e-9):6.1f} {a[4] / 1e6:9.1f} {a[5] / 1e6:9.1f}  {a[6]}/{a[7]}")
print(f"{'total (serialised, cold cache)':46s}     {tot:8.3f}")
if len(sys.argv) > 3:
    path, cfgi = sys.argv[2], sys.argv[3]
    try:
        tr = json.load(open(path))
    except Exception:
        tr = {"note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of one bench step, ncu on B200 (serialised launches); bench.py cites these as roofline.traffic"}
    align = sum(a[4] + a[5] for k, a in agg.items() if k.startswith("msa_") and "heavy" not in k)
    look = sum(a[4] + a[5] for k, a in agg.items() if "kmc_hash_lookup_kernel" in k)
    tr[f"config{cfgi}"] = {"align_pipeline": {"bytes": align, "source": f"{sys.argv[1].replace('gpurun_out', 'profiles')} (sum over the first-pass alignment launches of one step)"},
                           "kmc_hash_lookup_kernel": {"bytes": look, "source": f"{sys.argv[1].replace('gpurun_out', 'profiles')} (ncu, dram__bytes_read + dram__bytes_write of the launch)"}}
    json.dump(tr, open(path, "w"), indent=1)
    print("wrote", path)

#!/bin/bash
# round-2 GPU call c: random-access ceiling sweep; e2e host-thread sweep; per-lane contiguous arrays A/B
mkdir -p gpurun_out
python - > gpurun_out/r02c_gather.txt 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
from ploidyfrost_b200 import capi
ctx = capi.Context(0)
print("legacy pf_bench_random_gather (GB/s of 32-B sectors):", ctx.bench_random_gather(4 << 30))
for tab in (4 << 30, 32 << 30):
    for width in (32, 64, 16):
        for ilp in (4, 8, 16):
            for cps in (2, 4, 8):
                try:
                    v = ctx.bench_gather_sweep(tab, width, ilp, cps)
                    print(f"table {tab >> 30} GB width {width} ilp {ilp} ctas/sm {cps}: {v:.2f} G accesses/s = {v * width:.0f} GB/s", flush=True)
                except Exception as e:
                    print("ERR", tab, width, ilp, cps, e)
ctx.close()
PY
python bench.py --config 2 --steps 5 --warmup 3 --e2e-sweep 1,2,8 > gpurun_out/r02c_c2.json 2> gpurun_out/r02c_c2.err; echo "c2 rc=$?" > gpurun_out/r02c_rc.txt
PF_LANE_CONTIG=1 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu-baseline --e2e-threads 1 > gpurun_out/r02c_c2_contig.json 2> gpurun_out/r02c_c2_contig.err; echo "c2 contig rc=$?" >> gpurun_out/r02c_rc.txt
PF_LANE_CONTIG=1 python bench.py --config 1 --steps 5 --warmup 3 --e2e-threads 1 > gpurun_out/r02c_c1_contig.json 2> gpurun_out/r02c_c1_contig.err; echo "c1 contig rc=$?" >> gpurun_out/r02c_rc.txt
cat gpurun_out/r02c_rc.txt; grep -h "rank 0" gpurun_out/r02c_c*.err | cut -c1-200; tail -n 60 gpurun_out/r02c_gather.txt

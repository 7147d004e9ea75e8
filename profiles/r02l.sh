#!/bin/bash
# round-2 GPU call l: e2e leg timed over K steps in one bracket (host threads free-running); high-priority helper stream vs flat
mkdir -p gpurun_out
python bench.py --config 2 --steps 5 --warmup 3 --no-cpu-baseline --e2e-sweep 2x1,4x1,2x2,8x1,1x2 --e2e-profile gpurun_out/r02l_e2e_trace.json > gpurun_out/r02l_c2.json 2> gpurun_out/r02l_c2.err; echo "c2 rc=$?" > gpurun_out/r02l_rc.txt
python profiles/trace_summary.py gpurun_out/r02l_e2e_trace.json > gpurun_out/r02l_trace_summary.txt 2>&1
gzip -f gpurun_out/r02l_e2e_trace.json
PF_FLAT_PRIORITY=1 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu-baseline --e2e-sweep 4x1 > gpurun_out/r02l_c2_flat.json 2> gpurun_out/r02l_c2_flat.err; echo "flat rc=$?" >> gpurun_out/r02l_rc.txt
cat gpurun_out/r02l_rc.txt; cat gpurun_out/r02l_trace_summary.txt
python - <<'PY'
import json
for f in ("r02l_c2", "r02l_c2_flat"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"] / 1e6, 2), "step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "alone", d["e2e"].get("ms_per_step_synchronised_alone"), d["e2e"].get("ms_per_step_by_host_threads"))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY

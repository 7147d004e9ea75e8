#!/bin/bash
# round-2 GPU call k: CUPTI timeline of the e2e leg (where do 17 ms go when the device step is 11 ms?); configs[4] default batch
mkdir -p gpurun_out
python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline --e2e-profile gpurun_out/r02k_e2e_trace.json > gpurun_out/r02k_c2.json 2> gpurun_out/r02k_c2.err; echo "c2 rc=$?" > gpurun_out/r02k_rc.txt
python profiles/trace_summary.py gpurun_out/r02k_e2e_trace.json > gpurun_out/r02k_trace_summary.txt 2>&1
gzip -f gpurun_out/r02k_e2e_trace.json
python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/r02k_c4.json 2> gpurun_out/r02k_c4.err; echo "c4 rc=$?" >> gpurun_out/r02k_rc.txt
cat gpurun_out/r02k_rc.txt; cat gpurun_out/r02k_trace_summary.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02k_c4.json").read().strip().splitlines()[-1])
print("c4 value", d["value"], "ms", d["ms_per_step"], "align", d["ms_align_pipeline"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["config"]["batch_bubbles"], d.get("parity"))
PY

#!/bin/bash
mkdir -p gpurun_out
PF_BLOCKING_SYNC=1 timeout 95 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_sync_blocking.json 2> gpurun_out/r02_sync_blocking.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_sync_blocking.json").read().strip().splitlines()[-1])
print("blocking: step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "alone", round(d["e2e"]["ms_per_step_synchronised_alone"], 2))
PY

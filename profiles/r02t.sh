#!/bin/bash
# round-2 GPU call t: where does the estimation phase of the bound binary go? (release timing, with / without the warm-up batch)
mkdir -p gpurun_out
PF_PROGRAM_SKIP_REF=1 timeout 600 python integration/time_program.py 20000000 2 gpurun_out/r02t_warm.json > gpurun_out/r02t_warm.log 2>&1; echo "warm rc=$?" > gpurun_out/r02t_rc.txt
PF_NO_WARM=1 PF_PROGRAM_SKIP_REF=1 timeout 600 python integration/time_program.py 20000000 2 gpurun_out/r02t_nowarm.json > gpurun_out/r02t_nowarm.log 2>&1; echo "nowarm rc=$?" >> gpurun_out/r02t_rc.txt
cat gpurun_out/r02t_rc.txt
python - <<'PY'
import json
for f in ("r02t_warm", "r02t_nowarm"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    for k, v in d["runs"].items():
        print(f, k, {a: b for a, b in v.items() if a in ("wall_s", "estimation_phase_s", "phase_s", "open_wait_s", "collect_s", "device_wait_s", "release_s", "device_thread")})
PY

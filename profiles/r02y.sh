#!/bin/bash
# round-2 GPU call y: the final binaries -- whole programs (fresh box) then the integration / coloured tests
mkdir -p gpurun_out
timeout 300 python integration/time_program.py --colored 6000000 4 8 gpurun_out/r02y_prog_colored.json > gpurun_out/r02y_prog_colored.log 2>&1; echo "prog colored rc=$?" > gpurun_out/r02y_rc.txt
timeout 300 python integration/time_program.py 20000000 2 gpurun_out/r02y_prog_20m_dip.json > gpurun_out/r02y_prog_dip.log 2>&1; echo "prog dip rc=$?" >> gpurun_out/r02y_rc.txt
timeout 600 python -m pytest tests/test_gpu_integration.py tests/test_gpu_colored.py -x -q > gpurun_out/r02y_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02y_rc.txt; tail -3 gpurun_out/r02y_tests.log
cat gpurun_out/r02y_rc.txt
python - <<'PY'
import json
for f in ("r02y_prog_20m_dip", "r02y_prog_colored"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, json.dumps(d.get("summary")), d.get("tN_files_equal_as_multisets"))
    for k, v in d["runs"].items():
        if k.startswith("gpu -t 16"): print("   ", {a: b for a, b in v.items() if a in ("phase_s", "open_wait_s", "collect_s", "device_wait_s", "device_thread")})
PY

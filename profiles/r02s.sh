#!/bin/bash
# round-2 GPU call s: whole GPU suite with the new tests (coloured, host aligner path); whole programs again (warm-up batch) + coloured
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02s_tests.log 2>&1; echo "suite rc=$?" > gpurun_out/r02s_rc.txt
timeout 600 python integration/time_program.py 20000000 2 gpurun_out/r02s_prog_20m_dip.json > gpurun_out/r02s_prog_dip.log 2>&1; echo "prog dip rc=$?" >> gpurun_out/r02s_rc.txt
timeout 600 python integration/time_program.py --colored 6000000 4 8 gpurun_out/r02s_prog_colored.json > gpurun_out/r02s_prog_colored.log 2>&1; echo "prog colored rc=$?" >> gpurun_out/r02s_rc.txt
cat gpurun_out/r02s_rc.txt; tail -4 gpurun_out/r02s_tests.log
python - <<'PY'
import json
for f in ("r02s_prog_20m_dip", "r02s_prog_colored"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, json.dumps(d.get("summary")), d.get("tN_files_equal_as_multisets"))
        for k, v in d["runs"].items():
            print("   ", k, {a: b for a, b in v.items() if a in ("wall_s", "estimation_phase_s", "collect_s", "device_wait_s", "device_thread", "rc", "tail")})
    except Exception as e:
        print(f, "ERR", e)
PY

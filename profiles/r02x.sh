#!/bin/bash
# round-2 GPU call x: the final binaries (enlarged warm-up batch) through the integration + coloured tests and the whole-program timing; final bench line of record
mkdir -p gpurun_out
timeout 300 python integration/time_program.py 20000000 2 gpurun_out/r02x_prog_20m_dip.json > gpurun_out/r02x_prog_dip.log 2>&1; echo "prog dip rc=$?" > gpurun_out/r02x_rc.txt
timeout 300 python integration/time_program.py --colored 6000000 4 8 gpurun_out/r02x_prog_colored.json > gpurun_out/r02x_prog_colored.log 2>&1; echo "prog colored rc=$?" >> gpurun_out/r02x_rc.txt
timeout 900 python -m pytest tests/test_gpu_integration.py tests/test_gpu_colored.py tests/test_gpu_dropin.py -x -q > gpurun_out/r02x_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02x_rc.txt; tail -3 gpurun_out/r02x_tests.log
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02x_bench_c2.json 2> gpurun_out/r02x_bench_c2.err; echo "c2 rc=$?" >> gpurun_out/r02x_rc.txt
cat gpurun_out/r02x_rc.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02x_bench_c2.json").read().strip().splitlines()[-1])
print("c2 value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), d["e2e"].get("ms_per_step"), d["e2e"]["d2h_bytes_per_step"], d.get("parity"))
for f in ("r02x_prog_20m_dip", "r02x_prog_colored"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, json.dumps(d.get("summary")), d.get("tN_files_equal_as_multisets"))
PY

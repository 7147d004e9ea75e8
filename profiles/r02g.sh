#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_staged.py tests/test_gpu_align.py -x -q > gpurun_out/r02g_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r02g_rc.txt
python bench.py --config 2 --steps 5 --warmup 3 --e2e-sweep 1,2,8 > gpurun_out/r02g_c2.json 2> gpurun_out/r02g_c2.err; echo "c2 rc=$?" >> gpurun_out/r02g_rc.txt
python bench.py --config 2 --steps 5 --warmup 3 --no-staged-align --no-cpu-baseline > gpurun_out/r02g_c2_nostaged.json 2> gpurun_out/r02g_c2_nostaged.err; echo "c2 nostaged rc=$?" >> gpurun_out/r02g_rc.txt
cat gpurun_out/r02g_rc.txt; tail -4 gpurun_out/r02g_tests.log
python - <<'PY'
import json
for f in ("r02g_c2", "r02g_c2_nostaged"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"] / 1e6, 2), "M/s step", round(d["ms_per_step"], 2), "align", round(d["ms_align_pipeline"], 2), "e2e", json.dumps(d["e2e"]))
        print("   parity", d.get("parity"))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY

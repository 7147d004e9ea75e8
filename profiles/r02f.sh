#!/bin/bash
# round-2 GPU call f: SM-partition sweep (lookups beside the alignment), e2e threads x sub-batches sweep, new kmc tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kmc.py -x -q -k "seams or shared or cannot_search" > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r02f_rc.txt
python bench.py --config 2 --steps 5 --warmup 3 --lookup-sms-sweep 0,16,24,32,48,64 --e2e-sweep 4x1,4x4,8x2,2x4 --no-cpu-baseline > gpurun_out/r02f_c2.json 2> gpurun_out/r02f_c2.err; echo "c2 rc=$?" >> gpurun_out/r02f_rc.txt
PF_E2E=1 python bench.py --config 2 --steps 5 --warmup 3 --lookup-sms 32 --no-cpu-baseline > gpurun_out/r02f_c2_p32.json 2> gpurun_out/r02f_c2_p32.err; echo "c2 p32 rc=$?" >> gpurun_out/r02f_rc.txt
python bench.py --config 1 --steps 5 --warmup 3 --lookup-sms-sweep 0,16,32 --no-cpu-baseline --e2e-threads 4 > gpurun_out/r02f_c1.json 2> gpurun_out/r02f_c1.err; echo "c1 rc=$?" >> gpurun_out/r02f_rc.txt
cat gpurun_out/r02f_rc.txt; tail -3 gpurun_out/r02f_tests.log
python - <<'PY'
import json
for f in ("r02f_c2", "r02f_c2_p32", "r02f_c1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"] / 1e6, 2), "M/s step", round(d["ms_per_step"], 2), "lookup", round(d["ms_lookup_kernel"], 2), "align", round(d["ms_align_pipeline"], 2),
              "e2e ms", round(d["e2e"]["ms_per_step"], 2), d["e2e"].get("ms_per_step_by_host_threads"))
        print("   sweep", json.dumps(d.get("sm_partition_sweep")))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY

#!/bin/bash
# round-2 GPU call j: does the e2e leg suffer from stream -> hardware-queue aliasing (40 streams over the default 8 connections)?
mkdir -p gpurun_out
for mc in 8 32; do
  CUDA_DEVICE_MAX_CONNECTIONS=$mc python bench.py --config 2 --steps 5 --warmup 3 --no-cpu-baseline --e2e-sweep 1x1,2x2,8x1 > gpurun_out/r02j_c2_mc$mc.json 2> gpurun_out/r02j_c2_mc$mc.err; echo "mc$mc rc=$?" >> gpurun_out/r02j_rc.txt
done
PF_HEAVY_CTAS=0 CUDA_DEVICE_MAX_CONNECTIONS=32 python bench.py --config 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_c2_noheavy.json 2> gpurun_out/r02j_c2_noheavy.err; echo "noheavy rc=$?" >> gpurun_out/r02j_rc.txt
cat gpurun_out/r02j_rc.txt
python - <<'PY'
import json
for f in ("r02j_c2_mc8", "r02j_c2_mc32", "r02j_c2_noheavy"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"] / 1e6, 2), "step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), d["e2e"].get("ms_per_step_by_host_threads"))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY

#!/bin/bash
# round-2 GPU call w: A/B on one box -- offsets rebuilt on the host (default) against offsets copied from the device
mkdir -p gpurun_out
for i in 1 2; do
python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02w_host_$i.json 2> gpurun_out/r02w_host_$i.err
PF_COPY_OFFSETS=1 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02w_copy_$i.json 2> gpurun_out/r02w_copy_$i.err
done
python - <<'PY'
import json
for f in ("r02w_host_1", "r02w_copy_1", "r02w_host_2", "r02w_copy_2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "alone", round(d["e2e"]["ms_per_step_synchronised_alone"], 2), d["e2e"]["d2h_bytes_per_step"])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-600:])
PY

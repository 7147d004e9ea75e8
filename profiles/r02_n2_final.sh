#!/bin/bash
# the final commit under torchrun at N = 2 (sharded legs included, parity on a small sample)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --cpu-sample 8192 > gpurun_out/r02_n2_final.json 2> gpurun_out/r02_n2_final.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_n2_final.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["numa"])
print("sharded", json.dumps(d.get("sharded"))[:600]); print("parity", d.get("parity"))
PY

#!/bin/bash
# multi-GPU bench exactly as the driver launches it: bash profiles/r02_multi.sh N [extra bench args]
N=${1:-2}; shift
mkdir -p gpurun_out
python bench.py --impl reference --gpus $N --steps 3 --warmup 1 "$@" > gpurun_out/r02_ref_n$N.json 2> gpurun_out/r02_ref_n$N.err; echo "ref rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"
grep -h "rank" gpurun_out/r02_bench_n$N.err | cut -c1-220 | head -10; tail -c 1500 gpurun_out/r02_bench_n$N.err | tail -5
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n$N.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    print("sharded", json.dumps(d.get("sharded")))
    print("parity", d.get("parity"))
except Exception as e:
    print("no line:", e)
PY

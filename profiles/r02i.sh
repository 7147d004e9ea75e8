#!/bin/bash
# round-2 GPU call i: e2e call trace on configs[2]; launch list of the repo's own kernels for one step
mkdir -p gpurun_out
python bench.py --config 2 --steps 5 --warmup 3 --e2e-trace --no-cpu-baseline > gpurun_out/r02i_c2.json 2> gpurun_out/r02i_c2.err; echo "c2 rc=$?" > gpurun_out/r02i_rc.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02i_launches_c2.csv \
  --kernel-name 'regex:msa_|kmc_hash_lookup|site_|plan_|gather_|class_bounds|slot_size|collect_|reject_|result_size|cov_init|tile_seq|win_off|DeviceRadixSort|DeviceScan' \
  python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline --e2e-threads 1 > gpurun_out/r02i_ncu_launch.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r02i_rc.txt
cat gpurun_out/r02i_rc.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02i_c2.json").read().strip().splitlines()[-1])
print("value", round(d["value"] / 1e6, 2), "step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2))
print(json.dumps(d["e2e"].get("single_thread_call_ms")))
PY

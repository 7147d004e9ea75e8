#!/bin/bash
# round-2 GPU call n: the reference arm and the default bench exactly as the driver runs them at N = 1; whole programs on real graphs
mkdir -p gpurun_out
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02n_ref.json 2> gpurun_out/r02n_ref.err; echo "ref rc=$?" > gpurun_out/r02n_rc.txt
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench rc=$?" >> gpurun_out/r02n_rc.txt
PF_PROGRAM_CHECK=1 timeout 900 python integration/time_program.py 20000000 2 gpurun_out/r02n_prog_20m_dip.json > gpurun_out/r02n_prog_dip.log 2>&1; echo "prog dip rc=$?" >> gpurun_out/r02n_rc.txt
timeout 900 python integration/time_program.py 12000000 4 gpurun_out/r02n_prog_12m_tet.json > gpurun_out/r02n_prog_tet.log 2>&1; echo "prog tet rc=$?" >> gpurun_out/r02n_rc.txt
cat gpurun_out/r02n_rc.txt
python - <<'PY'
import json
for f in ("r02n_ref", "r02n_bench"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"), d.get("parity"))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
for f in ("r02n_prog_20m_dip", "r02n_prog_12m_tet"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, json.dumps(d.get("summary")), d.get("tN_files_equal_as_multisets"), d.get("t1_files_identical"), d.get("reference_tN_equals_its_own_second_run"), d.get("tN_differences"))
    except Exception as e:
        print(f, "ERR", e)
PY

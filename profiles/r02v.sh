#!/bin/bash
# round-2 GPU call v: whole programs FIRST on the fresh box (r02s / r02u measured them after minutes of benches and tests on the same
# box: the first lookup calls of the phase then take 0.5 - 1 s longer and the phase reads 1.2 - 1.6 s instead of 0.35 s; cause not identified);
# then the GPU suite and the bench line of record with the offsets rebuilt on the host
mkdir -p gpurun_out
free -g > gpurun_out/r02v_free_before.txt
timeout 600 python integration/time_program.py 20000000 2 gpurun_out/r02v_prog_20m_dip.json > gpurun_out/r02v_prog_dip.log 2>&1; echo "prog dip rc=$?" > gpurun_out/r02v_rc.txt
timeout 600 python integration/time_program.py 12000000 4 gpurun_out/r02v_prog_12m_tet.json > gpurun_out/r02v_prog_tet.log 2>&1; echo "prog tet rc=$?" >> gpurun_out/r02v_rc.txt
timeout 600 python integration/time_program.py --colored 6000000 4 8 gpurun_out/r02v_prog_colored.json > gpurun_out/r02v_prog_colored.log 2>&1; echo "prog colored rc=$?" >> gpurun_out/r02v_rc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02v_tests.log 2>&1; echo "suite rc=$?" >> gpurun_out/r02v_rc.txt; tail -3 gpurun_out/r02v_tests.log
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02v_bench_c2.json 2> gpurun_out/r02v_bench_c2.err; echo "c2 rc=$?" >> gpurun_out/r02v_rc.txt
free -g > gpurun_out/r02v_free_after.txt
cat gpurun_out/r02v_rc.txt
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02v_bench_c2.json").read().strip().splitlines()[-1])
    print("c2 value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), d["e2e"].get("ms_per_step"), d["e2e"]["d2h_bytes_per_step"], d.get("parity"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02v_bench_c2.err").read()[-800:])
for f in ("r02v_prog_20m_dip", "r02v_prog_12m_tet", "r02v_prog_colored"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, json.dumps(d.get("summary")), d.get("tN_files_equal_as_multisets"))
    except Exception as e:
        print(f, "ERR", e)
PY

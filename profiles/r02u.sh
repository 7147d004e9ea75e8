#!/bin/bash
# round-2 final GPU call: bench lines of record (configs[2] default with both arms, configs[1], configs[4]) and whole programs on real
# Bifrost graphs (20 Mbp diploid with the -t 1 byte check, 12 Mbp tetraploid, coloured 8 samples, 100 Mbp diploid)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02u_tests.log 2>&1; echo "suite rc=$?" > gpurun_out/r02u_rc0.txt; tail -3 gpurun_out/r02u_tests.log
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02u_ref.json 2> gpurun_out/r02u_ref.err; echo "ref rc=$?" > gpurun_out/r02u_rc.txt
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02u_bench_c2.json 2> gpurun_out/r02u_bench_c2.err; echo "c2 rc=$?" >> gpurun_out/r02u_rc.txt
python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/r02u_bench_c1.json 2> gpurun_out/r02u_bench_c1.err; echo "c1 rc=$?" >> gpurun_out/r02u_rc.txt
python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/r02u_bench_c4.json 2> gpurun_out/r02u_bench_c4.err; echo "c4 rc=$?" >> gpurun_out/r02u_rc.txt
PF_PROGRAM_CHECK=1 timeout 600 python integration/time_program.py 20000000 2 gpurun_out/r02u_prog_20m_dip.json > gpurun_out/r02u_prog_dip.log 2>&1; echo "prog dip rc=$?" >> gpurun_out/r02u_rc.txt
timeout 600 python integration/time_program.py 12000000 4 gpurun_out/r02u_prog_12m_tet.json > gpurun_out/r02u_prog_tet.log 2>&1; echo "prog tet rc=$?" >> gpurun_out/r02u_rc.txt
timeout 600 python integration/time_program.py --colored 6000000 4 8 gpurun_out/r02u_prog_colored.json > gpurun_out/r02u_prog_colored.log 2>&1; echo "prog colored rc=$?" >> gpurun_out/r02u_rc.txt
timeout 1000 python integration/time_program.py 100000000 2 gpurun_out/r02u_prog_100m_dip.json > gpurun_out/r02u_prog_100m.log 2>&1; echo "prog 100m rc=$?" >> gpurun_out/r02u_rc.txt
cat gpurun_out/r02u_rc.txt
python - <<'PY'
import json
for f in ("r02u_ref", "r02u_bench_c2", "r02u_bench_c1", "r02u_bench_c4"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), d["e2e"].get("ms_per_step"), "frac", (d.get("roofline") or {}).get("frac"), d.get("parity"))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-800:])
for f in ("r02u_prog_20m_dip", "r02u_prog_12m_tet", "r02u_prog_colored", "r02u_prog_100m_dip"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, json.dumps(d.get("summary")), d.get("tN_files_equal_as_multisets"), d.get("t1_files_identical"), d.get("reference_tN_equals_its_own_second_run"), d.get("tN_differences"))
    except Exception as e:
        print(f, "ERR", e)
PY

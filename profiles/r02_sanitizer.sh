#!/bin/bash
# compute-sanitizer over the smoke path and a small bench batch (memcheck + racecheck + synccheck); summaries -> profiles/
# run under gpurun: bash profiles/r02_sanitizer.sh
mkdir -p gpurun_out
export PF_HEAVY_CTAS=0        # the polling heavy-queue CTAs would wait on kernels a serialising tool never co-schedules
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_${tool}_smoke.log 2>&1
  echo "$tool smoke rc=$?" >> gpurun_out/r02_sanitizer_rc.txt
done
# a small three-phase batch of every kernel family: lane / group / CTA kernels, lookups, site k-mers
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python bench.py --config 4 --batch 96 --steps 1 --warmup 0 --no-cpu-baseline --e2e-threads 1 > gpurun_out/r02_sanitizer_memcheck_c4.log 2>&1
echo "memcheck c4 rc=$?" >> gpurun_out/r02_sanitizer_rc.txt
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python bench.py --config 4 --batch 48 --steps 1 --warmup 0 --no-cpu-baseline --e2e-threads 1 > gpurun_out/r02_sanitizer_racecheck_c4.log 2>&1
echo "racecheck c4 rc=$?" >> gpurun_out/r02_sanitizer_rc.txt
for f in gpurun_out/r02_sanitizer_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|smoke OK" $f | tail -5; done
cat gpurun_out/r02_sanitizer_rc.txt

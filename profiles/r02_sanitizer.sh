#!/bin/bash
# compute-sanitizer over the smoke path and a small three-phase batch of every kernel family (memcheck + racecheck + synccheck);
# summaries -> profiles/r02_sanitizer_summary.txt.  Run under gpurun: bash profiles/r02_sanitizer.sh
mkdir -p gpurun_out
rm -f gpurun_out/r02_sanitizer_rc.txt
export PF_HEAVY_CTAS=0        # the polling heavy-queue CTAs would wait on kernels a serialising tool never co-schedules
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_${tool}_smoke.log 2>&1
  echo "$tool smoke rc=$?" >> gpurun_out/r02_sanitizer_rc.txt
done
# lane / group / CTA kernels, hash lookups, site k-mers: configs[4] bubbles (50 bp .. 5 kbp, 2-4 rows) in a small batch
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python bench.py --config 4 --batch 64 --steps 1 --warmup 0 --no-cpu-baseline --e2e-threads 1 > gpurun_out/r02_sanitizer_memcheck_c4.log 2>&1
echo "memcheck c4 rc=$?" >> gpurun_out/r02_sanitizer_rc.txt
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python bench.py --config 4 --batch 32 --steps 1 --warmup 0 --no-cpu-baseline --e2e-threads 1 > gpurun_out/r02_sanitizer_racecheck_c4.log 2>&1
echo "racecheck c4 rc=$?" >> gpurun_out/r02_sanitizer_rc.txt
# the short-bubble kernels (lane s16x2, group) and the streaming open on a small tetraploid
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python bench.py --config 1 --genome-mbp 2 --batch 4096 --steps 1 --warmup 0 --no-cpu-baseline --e2e-threads 1 > gpurun_out/r02_sanitizer_memcheck_c1.log 2>&1
echo "memcheck c1 rc=$?" >> gpurun_out/r02_sanitizer_rc.txt
{
  for f in gpurun_out/r02_sanitizer_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|smoke OK|Error:|bench.py:" $f | sort | uniq -c | sort -rn | head -12; tail -2 $f | cut -c1-300; done
  cat gpurun_out/r02_sanitizer_rc.txt
} > gpurun_out/r02_sanitizer_summary.txt 2>&1
cat gpurun_out/r02_sanitizer_summary.txt

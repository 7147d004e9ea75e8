#!/bin/bash
# round-2 final profile call: launch list with per-launch metrics of one configs[2] step, ncu --set full of the dominant
# alignment launch (lane kernel, <= 64 class) and of the lookup kernel, with source correlation; summaries -> profiles/
mkdir -p gpurun_out
export PF_HEAVY_CTAS=0      # the polling heavy-queue CTAs would spin until their time-out under ncu's serialised launches
B="--config 2 --no-cpu-baseline --e2e-threads 1"
python bench.py $B --steps 1 --warmup 1 > gpurun_out/r02p_prime.json 2> gpurun_out/r02p_prime.err   # builds and caches the database (torch data tooling), outside ncu
K='regex:msa_|kmc_hash_lookup|site_|gather_kernel|plan_kernel|slot_size|result_size|collect_retry|reject_dash|tile_seq|cov_init|class_bounds|DeviceRadixSort|DeviceScan'
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_active.avg.per_cycle_active,launch__grid_size,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r02p_launches_c2.csv python bench.py $B --steps 1 --warmup 1 > gpurun_out/r02p_ncu_launch.log 2>&1
echo "launch list rc=$?" > gpurun_out/r02p_rc.txt
# lane kernel launches per step: <= 96 then <= 64 (heaviest first); one warm-up step first => the <= 64 launch of the timed step is #3
ncu --set full --clock-control none --import-source on -k regex:msa_lane_kernel -s 3 -c 1 -f -o gpurun_out/r02p_lane64_c2 python bench.py $B --steps 1 --warmup 1 > gpurun_out/r02p_ncu_lane64.log 2>&1
echo "lane64 rc=$?" >> gpurun_out/r02p_rc.txt
ncu --set full --clock-control none --import-source on -k regex:kmc_hash_lookup_kernel -s 1 -c 1 -f -o gpurun_out/r02p_lookup_c2 python bench.py $B --steps 1 --warmup 1 > gpurun_out/r02p_ncu_lookup.log 2>&1
echo "lookup rc=$?" >> gpurun_out/r02p_rc.txt
python profiles/ncu_lines.py gpurun_out/r02p_lane64_c2.ncu-rep msa_lane_kernelILi2E > gpurun_out/r02p_lane64_lines.txt 2>&1
for k in lane64 lookup; do
  ncu -i gpurun_out/r02p_${k}_c2.ncu-rep --page details --csv > gpurun_out/r02p_${k}_details.csv 2>/dev/null
  ncu -i gpurun_out/r02p_${k}_c2.ncu-rep --page details > gpurun_out/r02p_${k}_ncu.txt 2>/dev/null
done
python profiles/launch_metrics.py gpurun_out/r02p_launches_c2.csv > gpurun_out/r02p_launch_metrics.txt 2>&1
cat gpurun_out/r02p_rc.txt; cat gpurun_out/r02p_launch_metrics.txt | head -30

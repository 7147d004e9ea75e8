#!/bin/bash
# round-2 GPU call e: streaming open parity (kmc + sharded tests), launch list + per-launch metrics on configs[2], full ncu capture of
# the two lane-kernel launches and the lookup kernel with source correlation
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kmc.py tests/test_gpu_sharded.py tests/test_gpu_e2e.py -x -q > gpurun_out/r02e_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r02e_rc.txt
export PF_HEAVY_CTAS=0
B="--config 2 --no-cpu-baseline --e2e-threads 1"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_active.avg.per_cycle_active,launch__grid_size,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:'msa_|kmc_|site_|gather_kernel|plan_kernel|slot_size|result_size|collect_retry|reject_dash|DeviceRadixSort|DeviceScan|tile_seq|cov_init' -c 200 --csv --log-file gpurun_out/r02e_launches_c2.csv \
    python bench.py $B --steps 1 --warmup 1 > gpurun_out/r02e_ncu_launch.log 2>&1
# lane kernel launches per step: <=96 then <=64 (heaviest first); warm-up step first => skip 2 / 3
ncu --set full --clock-control none --import-source on -k regex:msa_lane_kernel -s 3 -c 1 -f -o gpurun_out/r02e_lane64_c2 python bench.py $B --steps 1 --warmup 1 > gpurun_out/r02e_ncu_lane64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kmc_hash_lookup_kernel -s 1 -c 1 -f -o gpurun_out/r02e_lookup_c2 python bench.py $B --steps 1 --warmup 1 > gpurun_out/r02e_ncu_lookup.log 2>&1
python profiles/ncu_lines.py gpurun_out/r02e_lane64_c2.ncu-rep msa_lane_kernelILi2E > gpurun_out/r02e_lane64_lines.txt 2>&1
ncu -i gpurun_out/r02e_lane64_c2.ncu-rep --page details --csv > gpurun_out/r02e_lane64_details.csv 2>/dev/null
ncu -i gpurun_out/r02e_lookup_c2.ncu-rep --page details --csv > gpurun_out/r02e_lookup_details.csv 2>/dev/null
PF_PROGRAM_SKIP_REF=1 timeout 600 python integration/time_program.py 20000000 2 gpurun_out/r02e_prog_20m_dip.json > gpurun_out/r02e_prog.log 2>&1; echo "prog rc=$?" >> gpurun_out/r02e_rc.txt
grep -o 'device_thread[^}]*' gpurun_out/r02e_prog.log | head -3
cat gpurun_out/r02e_rc.txt; tail -3 gpurun_out/r02e_tests.log; head -50 gpurun_out/r02e_lane64_lines.txt

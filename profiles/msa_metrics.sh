#!/bin/bash
# per-launch metrics of the alignment kernels of one bench step (serialised by ncu): profiles/msa_metrics.sh TAG
TAG=${1:-x}
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_active.avg.per_cycle_active,launch__grid_size,launch__shared_mem_per_block_dynamic,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:msa_ -c ${NCAP:-8} --csv --log-file gpurun_out/msa_metrics_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline $BENCH_ARGS > gpurun_out/m_$TAG.log 2>&1

#!/bin/bash
# profiles/run_ncu.sh TAG "KERNEL_REGEXES" [bench args...] -- run under gpurun (1 GPU).  Writes into gpurun_out/:
#   launches_TAG.csv   every launch of OUR kernels in a short bench run with its device time (cold-cache, serialised:
#                      shares only); torch's data-generation kernels are filtered out by name
#   K_TAG.ncu-rep      ncu --set full of launch #SKIP of each kernel regex K (default: second launch = after warm-up)
TAG=${1:-r01}; KRN=${2:-msa_lane_kernel}; shift; shift
ARGS=${@:---batch 262144}
OURS='regex:msa_|kmc_|tile_seq|reject_dash|plan_kernel|class_bounds|slot_size|collect_retry|result_size|gather_kernel|cov_init|repack_|DeviceRadixSort|DeviceScan'
mkdir -p gpurun_out
[ -n "$NO_LAUNCH_LIST" ] || ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py $ARGS --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
for K in $KRN; do
ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-1} -c 1 -f -o gpurun_out/${K}_$TAG \
    python bench.py $ARGS --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
ls -la gpurun_out/ | tail -5

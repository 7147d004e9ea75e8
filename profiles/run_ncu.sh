#!/bin/bash
# profiles/run_ncu.sh TAG KERNEL_REGEX [bench args...] -- run under gpurun (1 GPU).  Writes into gpurun_out/:
#   launches_TAG.csv   every launch of a short bench run with its device time (cold-cache, serialised: shares only)
#   KERNEL_TAG.ncu-rep ncu --set full of the first launch matching KERNEL_REGEX after the warm-up step
TAG=${1:-r01}; KRN=${2:-msa_lane_kernel}; shift; shift
ARGS=${@:---genome-mbp 5 --batch 65536}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py $ARGS --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
for K in $KRN; do
ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -f -o gpurun_out/${K}_$TAG \
    python bench.py $ARGS --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
ls -la gpurun_out/

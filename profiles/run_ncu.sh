#!/bin/bash
# profiles/run_ncu.sh TAG [bench args...] -- run under gpurun (1 GPU).  Writes into gpurun_out/:
#   launches_TAG.csv   every launch of a short bench run with its device time (cold-cache, serialised: shares only)
#   msa_TAG.ncu-rep    ncu --set full of the alignment kernel;  lookup_TAG.ncu-rep  same for the KMC lookup kernel
TAG=${1:-r01}; shift
ARGS=${@:---genome-mbp 5 --batch 65536}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py $ARGS --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:msa_kernel -s 1 -c 1 -f -o gpurun_out/msa_$TAG \
    python bench.py $ARGS --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_msa_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kmc_lookup_kernel -s 1 -c 1 -f -o gpurun_out/lookup_$TAG \
    python bench.py $ARGS --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_lookup_$TAG.log 2>&1
ls -la gpurun_out/

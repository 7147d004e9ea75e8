#!/bin/bash
# round-2 first GPU call: what the box is, how long the configs[2] database takes to build/open, first bench lines
mkdir -p gpurun_out
{ nproc; free -g; df -h /tmp /dev/shm; nvidia-smi --query-gpu=name,memory.total --format=csv; } > gpurun_out/r02_box.txt 2>&1
python bench.py --config 2 --steps 5 --warmup 3 > gpurun_out/r02a_c2.json 2> gpurun_out/r02a_c2.err; echo "c2 rc=$?" >> gpurun_out/r02_box.txt
python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/r02a_c1.json 2> gpurun_out/r02a_c1.err; echo "c1 rc=$?" >> gpurun_out/r02_box.txt
timeout 600 python bench.py --config 4 --batch 512 --steps 2 --warmup 1 > gpurun_out/r02a_c4.json 2> gpurun_out/r02a_c4.err; echo "c4 rc=$?" >> gpurun_out/r02_box.txt
timeout 900 python -m pytest tests/test_gpu_differential.py -x -q > gpurun_out/r02a_diff.log 2>&1; echo "diff rc=$?" >> gpurun_out/r02_box.txt
tail -3 gpurun_out/r02a_diff.log; cat gpurun_out/r02_box.txt; tail -c 600 gpurun_out/r02a_c2.err

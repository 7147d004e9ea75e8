#!/bin/bash
# round-2 GPU call b: new traceback + CTA kernel -- parity tests, then the three bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_align.py tests/test_gpu_differential.py -x -q > gpurun_out/r02b_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r02b_rc.txt
timeout 600 python bench.py --config 4 --batch 512 --steps 3 --warmup 2 > gpurun_out/r02b_c4.json 2> gpurun_out/r02b_c4.err; echo "c4 rc=$?" >> gpurun_out/r02b_rc.txt
python bench.py --config 2 --steps 5 --warmup 3 > gpurun_out/r02b_c2.json 2> gpurun_out/r02b_c2.err; echo "c2 rc=$?" >> gpurun_out/r02b_rc.txt
python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/r02b_c1.json 2> gpurun_out/r02b_c1.err; echo "c1 rc=$?" >> gpurun_out/r02b_rc.txt
tail -15 gpurun_out/r02b_tests.log; cat gpurun_out/r02b_rc.txt; grep -h "rank 0" gpurun_out/r02b_c*.err | cut -c1-300

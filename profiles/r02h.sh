#!/bin/bash
# round-2 GPU call h: the coloured drop-in against the unmodified binary, whole GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_colored.py -x -q > gpurun_out/r02h_colored.log 2>&1; echo "colored rc=$?" > gpurun_out/r02h_rc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_tests.log 2>&1; echo "suite rc=$?" >> gpurun_out/r02h_rc.txt
cat gpurun_out/r02h_rc.txt; tail -5 gpurun_out/r02h_colored.log; tail -5 gpurun_out/r02h_tests.log

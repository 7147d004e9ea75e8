#!/bin/bash
mkdir -p gpurun_out
timeout 160 python -m pytest tests/test_gpu_integration.py tests/test_gpu_colored.py -x -q > gpurun_out/r02_last_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_last_tests.log

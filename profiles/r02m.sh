#!/bin/bash
# round-2 GPU call m (2 GPUs): the bench exactly as the driver launches it at N = 2 (reference arm, then torchrun), sharded legs included
bash profiles/r02_multi.sh 2 > gpurun_out/r02m_multi.log 2>&1
cat gpurun_out/r02m_multi.log | tail -30

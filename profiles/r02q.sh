#!/bin/bash
# round-2 GPU call q (8 GPUs): the bench exactly as the driver launches it at N = 8 (reference arm, then torchrun), sharded legs included
bash profiles/r02_multi.sh 8 > gpurun_out/r02q_multi.log 2>&1
tail -12 gpurun_out/r02q_multi.log | cut -c1-1200

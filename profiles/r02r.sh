#!/bin/bash
# round-2 GPU call r (8 GPUs): e2e with every rank bound to the NUMA node of its GPU, against unbound; no CPU legs (short call)
mkdir -p gpurun_out
nproc > gpurun_out/r02r_host.txt; lscpu | grep -i "numa\|socket\|model name" >> gpurun_out/r02r_host.txt; nvidia-smi topo -m >> gpurun_out/r02r_host.txt 2>&1
for tag in bound unbound; do
  extra=""; [ $tag = unbound ] && extra="--no-numa-bind"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-sharded-legs $extra > gpurun_out/r02r_n8_$tag.json 2> gpurun_out/r02r_n8_$tag.err; echo "$tag rc=$?"
  grep -h "e2e" gpurun_out/r02r_n8_$tag.err | grep -o "rank [0-9]\] step [0-9.]* ms.*e2e [0-9.]* ms ([^)]*)" | sort -u
done
head -40 gpurun_out/r02r_host.txt

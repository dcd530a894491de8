#!/bin/bash
# 4-GPU data-parallel sanity run (NCCL all-reduces inside the CUDA graph, clean exit)
mkdir -p gpurun_out
S=$(date +%s)
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 10 --warmup 3 --no-micro --no-cpu-baseline > gpurun_out/r2c30_dp4.json 2> gpurun_out/r2c30_dp4.err; rc=$?
echo "dp4 rc=$rc after $(( $(date +%s) - S )) s"; tail -1 gpurun_out/r2c30_dp4.json | cut -c1-260; grep -i "error\|capture" gpurun_out/r2c30_dp4.err | head -3 | cut -c1-200

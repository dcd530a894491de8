#!/bin/bash
# 2-GPU: does the data-parallel bench exit cleanly now (graph with captured NCCL released before process-group teardown)?
mkdir -p gpurun_out
S=$(date +%s)
timeout -k 10 360 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 --no-micro --no-cpu-baseline > gpurun_out/r2c21_dp2.json 2> gpurun_out/r2c21_dp2.err; rc=$?
echo "dp2 rc=$rc after $(( $(date +%s) - S )) s"; tail -1 gpurun_out/r2c21_dp2.json | cut -c1-300; tail -2 gpurun_out/r2c21_dp2.err | cut -c1-200
timeout -k 5 60 python tools/text_attn_bench.py 2>&1 | tail -1

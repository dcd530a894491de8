#!/bin/bash
# 2-GPU data-parallel bench: whole step incl. NCCL all-reduces captured in the CUDA graph + head-gradient overlap,
# versus the round-1 behaviour (compute-only graph, one exposed all-reduce)
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 10 --warmup 3 --no-micro --no-cpu-baseline; }
timeout 400 bash -c "$(declare -f run); run 29511" > gpurun_out/r2c9_dp2.json 2> gpurun_out/r2c9_dp2.err; echo "dp2 rc=$?"; tail -1 gpurun_out/r2c9_dp2.json | cut -c1-700; tail -3 gpurun_out/r2c9_dp2.err | cut -c1-300
E4T_GRAPH_NCCL=0 timeout 400 bash -c "$(declare -f run); run 29512" > gpurun_out/r2c9_dp2_nographnccl.json 2> gpurun_out/r2c9_dp2_nographnccl.err; echo "dp2 (eager exchange) rc=$?"; tail -1 gpurun_out/r2c9_dp2_nographnccl.json | cut -c1-400
timeout 300 python bench.py --steps 10 --warmup 3 --no-micro --no-cpu-baseline > gpurun_out/r2c9_dp1.json 2>/dev/null; echo "dp1 rc=$?"; cut -c1-200 gpurun_out/r2c9_dp1.json

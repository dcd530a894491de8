#!/bin/bash
# 2-GPU: WeightOffsets factor exchange against the slice all-reduce (same seeds, 2 eager steps), then its timing
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
E4T_WO_FACTOR_EXCHANGE=1 timeout -k 10 150 bash -c "$(declare -f run); run 29541 tools/dp_wo_exchange_check.py gpurun_out/r2c31_f1.pt" > gpurun_out/r2c31_f1.log 2>&1; echo "factor rc=$?"
E4T_WO_FACTOR_EXCHANGE=0 timeout -k 10 150 bash -c "$(declare -f run); run 29542 tools/dp_wo_exchange_check.py gpurun_out/r2c31_f0.pt" > gpurun_out/r2c31_f0.log 2>&1; echo "slice rc=$?"
python tools/dp_wo_exchange_check.py --compare gpurun_out/r2c31_f1.pt gpurun_out/r2c31_f0.pt 2>&1 | tail -3; tail -2 gpurun_out/r2c31_f1.log | cut -c1-200
timeout -k 10 200 bash -c "$(declare -f run); run 29543 bench.py --gpus 2 --steps 10 --warmup 3 --no-micro --no-cpu-baseline" 2>/dev/null | tail -1 | cut -c1-230
rm -f gpurun_out/r2c31_f1.pt gpurun_out/r2c31_f0.pt

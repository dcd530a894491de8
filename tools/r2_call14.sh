#!/bin/bash
# 1-GPU, every step under its own short timeout: lean single-thread role loops in fwd2 / fwd3 / fused backward
mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests/test_variants_gpu.py tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2c14_pytest.log; rc=${PIPESTATUS[0]}; echo "pytest rc=$rc"; cut -c1-300 gpurun_out/r2c14_pytest.log
if [ "$rc" = "0" ]; then
  timeout -k 5 200 python tools/attn_bench.py fwd bwd > gpurun_out/r2c14_attn.log 2>&1; echo "attn_bench rc=$?"; cut -c1-1100 gpurun_out/r2c14_attn.log
  timeout -k 5 200 python bench.py --steps 10 --warmup 3 --no-micro --no-cpu-baseline 2>/dev/null | cut -c1-160
  E4T_ATTN_FWD2=d timeout -k 5 200 python bench.py --steps 10 --warmup 3 --no-micro --no-cpu-baseline 2>/dev/null | cut -c1-160
  E4T_ATTN_FWD2=d timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:attn_fwd3 -c 1 -o gpurun_out/r2c14_fwd3 python tools/attn_bench.py fwd one > gpurun_out/r2c14_ncu.log 2>&1; tail -2 gpurun_out/r2c14_ncu.log
  timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_fused -c 1 -o gpurun_out/r2c14_bwd python tools/attn_bench.py bwd one > gpurun_out/r2c14_ncu_bwd.log 2>&1; tail -2 gpurun_out/r2c14_ncu_bwd.log
fi

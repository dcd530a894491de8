#!/bin/bash
# 1-GPU, every step under its own short timeout: reordered backward issue (double-buffered dS^T), fwd3 default + exp2 splits,
# one ncu capture of the K=320 GEMM (epilogue-bound shape)
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_variants_gpu.py tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2c15_pytest.log; rc=${PIPESTATUS[0]}; echo "pytest rc=$rc"; cut -c1-300 gpurun_out/r2c15_pytest.log
if [ "$rc" = "0" ]; then
  timeout -k 5 240 python tools/attn_bench.py fwd bwd > gpurun_out/r2c15_attn.log 2>&1; echo "attn_bench rc=$?"; cut -c1-1500 gpurun_out/r2c15_attn.log | head -3
  E4T_ATTN_BWD_REORD=0 timeout -k 5 100 python tools/attn_bench.py bwd one 2>&1 | cut -c1-300
  timeout -k 5 200 python bench.py --steps 10 --warmup 3 --no-micro --no-cpu-baseline 2>/dev/null | cut -c1-160
  timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_fused -c 1 -o gpurun_out/r2c15_bwd python tools/attn_bench.py bwd one > gpurun_out/r2c15_ncu_bwd.log 2>&1; tail -1 gpurun_out/r2c15_ncu_bwd.log
fi
timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:e4t_gemm --launch-skip 4 -c 1 -o gpurun_out/r2c15_gemm python tools/prof_gemm.py > gpurun_out/r2c15_ncu_gemm.log 2>&1; tail -1 gpurun_out/r2c15_ncu_gemm.log

#!/bin/bash
# 1-GPU: parity of the new kernels (double-buffered forward "d", per-warp GEMM epilogue), their timings, one ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_variants_gpu.py tests/test_tuning_gpu.py tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c10_pytest.log; tail -6 gpurun_out/r2c10_pytest.log
timeout 300 python tools/attn_bench.py fwd > gpurun_out/r2c10_attn_fwd.log 2>&1; tail -30 gpurun_out/r2c10_attn_fwd.log
timeout 200 python tools/gemm_epi_compare.py > gpurun_out/r2c10_gemm_epi.log 2>&1; cat gpurun_out/r2c10_gemm_epi.log
for m in 1 2; do E4T_GEMM_EPI_PLAIN=$m timeout 200 python bench.py --steps 10 --warmup 3 --no-micro --no-cpu-baseline 2>/dev/null | cut -c1-160; done
E4T_ATTN_FWD2=d timeout 200 python bench.py --steps 10 --warmup 3 --no-micro --no-cpu-baseline 2>/dev/null | cut -c1-160
E4T_ATTN_FWD2=d timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd3 -c 1 -o gpurun_out/r2c10_fwd3 python tools/attn_bench.py fwd one > gpurun_out/r2c10_ncu.log 2>&1; tail -3 gpurun_out/r2c10_ncu.log

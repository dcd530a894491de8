#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 60 tools/micro/launch_gap > gpurun_out/r2c19_launch_gap.log 2>&1; echo "launch_gap rc=$?"; cat gpurun_out/r2c19_launch_gap.log
timeout -k 5 300 python -m pytest tests/test_variants_gpu.py tests/test_kernels_gpu.py tests/test_gemm_gpu.py tests/test_tuning_gpu.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2c19_pytest.log; rc=${PIPESTATUS[0]}; echo "pytest rc=$rc"; cut -c1-300 gpurun_out/r2c19_pytest.log
if [ "$rc" = "0" ]; then
  for r in 0 1 2; do echo "E4T_ATTN_BWD_REORD=$r"; E4T_ATTN_BWD_REORD=$r timeout -k 5 100 python tools/attn_bench.py bwd one 2>&1 | cut -c90-200; done
  timeout -k 5 200 python bench.py --steps 10 --warmup 3 --no-micro --no-cpu-baseline 2>/dev/null | cut -c1-160
  E4T_ATTN_BWD_REORD=2 timeout -k 5 200 python bench.py --steps 10 --warmup 3 --no-micro --no-cpu-baseline 2>/dev/null | cut -c1-160
fi

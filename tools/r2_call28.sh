#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_variants_gpu.py tests/test_e2e_gpu.py -m gpu -x -q 2>&1 | tail -3
for p in 1 0 1 0; do echo "E4T_GEMM_EPI_PLAIN=$p"; E4T_GEMM_EPI_PLAIN=$p timeout -k 5 200 python bench.py --steps 15 --warmup 3 --no-micro --no-cpu-baseline 2>/dev/null | cut -c1-160; done

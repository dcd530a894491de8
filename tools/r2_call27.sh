#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_variants_gpu.py tests/test_e2e_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout -k 5 120 python tools/gemm_res_bench.py 2>&1 | tail -8
for i in 1 2; do timeout -k 5 200 python bench.py --steps 10 --warmup 3 --no-micro --no-cpu-baseline 2>/dev/null | cut -c1-160; done

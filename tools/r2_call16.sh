#!/bin/bash
# 1-GPU, every step under its own timeout: backward with prefetched row statistics (one barrier per block), fwd3 p2 default;
# then the whole GPU suite
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_variants_gpu.py tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2c16_pytest.log; rc=${PIPESTATUS[0]}; echo "pytest rc=$rc"; cut -c1-300 gpurun_out/r2c16_pytest.log
if [ "$rc" = "0" ]; then
  timeout -k 5 100 python tools/attn_bench.py bwd fwd one 2>&1 | cut -c1-1200
  timeout -k 5 200 python bench.py --steps 10 --warmup 3 --no-micro --no-cpu-baseline 2>/dev/null | cut -c1-160
  timeout -k 5 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2c16_pytest_all.log; echo "full suite rc=${PIPESTATUS[0]}"; cut -c1-300 gpurun_out/r2c16_pytest_all.log
fi

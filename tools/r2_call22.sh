#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests/test_tuning_gpu.py -m gpu -x -q -k "attention" 2>&1 | tail -6 > gpurun_out/r2c22_pytest.log; rc=${PIPESTATUS[0]}; echo "pytest rc=$rc"; cut -c1-400 gpurun_out/r2c22_pytest.log
if [ "$rc" = "0" ]; then
  timeout -k 5 300 python -m pytest tests/test_e2e_gpu.py tests/test_variants_gpu.py -m gpu -x -q 2>&1 | tail -4
  timeout -k 5 200 python bench.py --steps 10 --warmup 3 --no-micro --no-cpu-baseline 2>/dev/null | cut -c1-160
fi

#!/bin/bash
# 1-GPU, every step under its own short timeout: fwd3 ("d") localisation + parity, then timings / ncu only if parity passed
mkdir -p gpurun_out
timeout -k 5 120 python tools/diag_fwd3.py > gpurun_out/r2c13_diag.log 2>&1; echo "diag rc=$?"; cut -c1-200 gpurun_out/r2c13_diag.log
timeout -k 5 150 python -m pytest tests/test_variants_gpu.py -m gpu -x -q -k "forward_variants and -d]" 2>&1 | tail -8 > gpurun_out/r2c13_pytest_d.log; rc=${PIPESTATUS[0]}; echo "pytest d rc=$rc"; cut -c1-300 gpurun_out/r2c13_pytest_d.log
if [ "$rc" = "0" ]; then
  timeout -k 5 150 python tools/attn_bench.py fwd > gpurun_out/r2c13_attn_fwd.log 2>&1; echo "attn_bench rc=$?"; cut -c1-900 gpurun_out/r2c13_attn_fwd.log
  E4T_ATTN_FWD2=d timeout -k 5 200 python bench.py --steps 10 --warmup 3 --no-micro --no-cpu-baseline 2>/dev/null | cut -c1-160
  E4T_ATTN_FWD2=d timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:attn_fwd3 -c 1 -o gpurun_out/r2c13_fwd3 python tools/attn_bench.py fwd one > gpurun_out/r2c13_ncu.log 2>&1; tail -2 gpurun_out/r2c13_ncu.log
fi

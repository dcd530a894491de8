#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 400 python tools/sweep_r2.py gemm > gpurun_out/r2c20_gemm_sweep.log 2>&1; echo "sweep rc=$?"; head -3 gpurun_out/r2c20_gemm_sweep.log | cut -c1-300

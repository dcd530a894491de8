#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 120 python tools/diag_fwd3.py > gpurun_out/r2c12_diag.log 2>&1; echo "rc=$?"; cat gpurun_out/r2c12_diag.log | cut -c1-400

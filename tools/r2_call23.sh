#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python tools/glue_probe.py > gpurun_out/r2c23_glue.log 2>&1; echo "rc=$?"; tail -50 gpurun_out/r2c23_glue.log | cut -c1-330

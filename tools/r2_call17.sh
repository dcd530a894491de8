#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 60 tools/micro/launch_gap > gpurun_out/r2c17_launch_gap.log 2>&1; echo "launch_gap rc=$?"; cat gpurun_out/r2c17_launch_gap.log
for cg in "4,4,4,1" "4,4,4,2"; do echo "E4T_ATTN_CG=$cg"; E4T_ATTN_CG=$cg timeout -k 5 120 python tools/attn_bench.py fwd 2>&1 | grep "cross\|clip" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['shape'], d['fwd[0]']['ms'], d['fwd[d]']['ms'])"; done
E4T_ATTN_CG=4,4,4,2 timeout -k 5 120 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attn or attention" 2>&1 | tail -3
timeout -k 5 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2c17_pytest_all.log; echo "full suite rc=${PIPESTATUS[0]}"; cut -c1-300 gpurun_out/r2c17_pytest_all.log

#!/bin/bash
# one gpurun call: sweep the ping-pong variants, then (if one wins) the full GPU suite + bench under that variant
mkdir -p gpurun_out
timeout 330 python tools/pp_sweep.py > gpurun_out/pp_sweep.log 2>&1
BEST=$(cat gpurun_out/pp_best.txt 2>/dev/null || echo 0)
echo "BEST=$BEST"
cat gpurun_out/pp_sweep.log | tail -5
if [ "$BEST" != "0" ]; then
  E4T_ATTN_PP=$BEST timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pp_tests.log 2>&1; echo "tests rc=$?"
  tail -3 gpurun_out/pp_tests.log
  E4T_ATTN_PP=$BEST timeout 240 python bench.py > gpurun_out/pp_bench.json 2> gpurun_out/pp_bench.err; echo "bench rc=$?"
  cat gpurun_out/pp_bench.json
fi

#!/bin/bash
# Suggested FIRST gpurun call of the next round (≈3-4 min of GPU time):
#   gpurun --timeout 420 -- 'bash tools/r2_first_call.sh'
# 1. parity of every opt-in variant against the default kernels (per-test timeout: a deadlocked variant cannot eat the call)
# 2. one-process timing sweep of the variants, GroupNorm launch configs and every GEMM/conv signature of the step
# 3. ncu --set full --import-source captures of the fused attention backward: default, ping-pong, dQ-TMA
mkdir -p gpurun_out
E4T_TEST_OPTIN=1 timeout 150 python -m pytest tests/test_optin_gpu.py -m gpu -q --timeout 40 -p no:cacheprovider \
    > gpurun_out/r2_optin_tests.log 2>&1
echo "optin tests rc=$?"; tail -15 gpurun_out/r2_optin_tests.log
timeout 200 python tools/sweep_r2.py attn norm gemm > gpurun_out/r2_sweep.log 2>&1
echo "sweep rc=$?"; grep -E "^\[(attn|norm|gemm)\]" gpurun_out/r2_sweep.log | cut -c1-330
for v in "default" "E4T_ATTN_PP=4" "E4T_ATTN_DQ_TMA=1" "E4T_ATTN_PT_TMEM=1"; do
  name=$(echo "$v" | tr '=' '_')
  if [ "$v" = "default" ]; then envs=""; else envs="$v"; fi
  env $envs timeout 90 ncu --set full --import-source on --clock-control none -k regex:attn_bwd_fused -c 1 \
      -o gpurun_out/r2_bwd_$name -f python tools/prof_attn.py > gpurun_out/r2_ncu_$name.log 2>&1
  echo "ncu $v rc=$?"
done

#!/bin/bash
# Round-2 evidence run: whole GPU suite, smoke(), bench (default / unfreeze / tuning), ncu launch list of one step,
# ncu --set full of the shipped attention forward and backward kernels
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2c24_tests.log 2>&1; echo "all gpu tests rc=$?"; grep -E "^\[|passed|failed|^FAILED|^E  " gpurun_out/r2c24_tests.log | cut -c1-330 | tail -24
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout -k 5 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c24_bench.json 2> gpurun_out/r2c24_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2c24_bench.json; tail -2 gpurun_out/r2c24_bench.err | cut -c1-200
timeout -k 5 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-micro --unfreeze-clip-vision > gpurun_out/r2c24_bench_unfreeze.json 2>/dev/null; echo "unfreeze rc=$?"; cut -c1-200 gpurun_out/r2c24_bench_unfreeze.json
timeout -k 5 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-micro --tuning > gpurun_out/r2c24_bench_tuning.json 2>/dev/null; echo "tuning rc=$?"; cut -c1-200 gpurun_out/r2c24_bench_tuning.json
timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:attn_fwd3 -c 1 -o gpurun_out/r2c24_fwd3 python tools/attn_bench.py fwd one > gpurun_out/r2c24_ncu_fwd.log 2>&1; tail -1 gpurun_out/r2c24_ncu_fwd.log
timeout -k 5 200 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_fused -c 1 -o gpurun_out/r2c24_bwd python tools/attn_bench.py bwd one > gpurun_out/r2c24_ncu_bwd.log 2>&1; tail -1 gpurun_out/r2c24_ncu_bwd.log
timeout -k 5 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02b.csv python bench.py --profile-one-step --warmup 1 --no-cpu-baseline > gpurun_out/r2c24_launches.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/launches_r02b.csv

#!/bin/bash
# final round-2 evidence refresh on the shipped code
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2c29_tests.log 2>&1; echo "all gpu tests rc=$?"; grep -E "passed|failed|^FAILED" gpurun_out/r2c29_tests.log | tail -3
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -k 5 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c29_bench.json 2> gpurun_out/r2c29_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2c29_bench.json
timeout -k 5 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02c.csv python bench.py --profile-one-step --warmup 1 --no-cpu-baseline > gpurun_out/r2c29_launches.log 2>&1; echo "launch list rc=$?"

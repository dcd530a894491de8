#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s --timeout 400 -p no:cacheprovider > gpurun_out/r2c8_tests.log 2>&1; echo "all gpu tests rc=$?"; grep -E "^\[|passed|failed|^FAILED|^E  " gpurun_out/r2c8_tests.log | cut -c1-330 | tail -30
timeout 100 python tools/gemm_probe.py > gpurun_out/r2c8_gemm_probe.log 2>&1; echo "gemm_probe rc=$?"; grep "BN=256" gpurun_out/r2c8_gemm_probe.log | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2c8_bench.json; tail -2 gpurun_out/r2c8_bench.err
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-micro --unfreeze-clip-vision > gpurun_out/r2c8_bench_unfreeze.json 2> gpurun_out/r2c8_bench_unfreeze.err; echo "bench unfreeze rc=$?"; cut -c1-260 gpurun_out/r2c8_bench_unfreeze.json; tail -2 gpurun_out/r2c8_bench_unfreeze.err
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-micro --tuning > gpurun_out/r2c8_bench_tuning.json 2> gpurun_out/r2c8_bench_tuning.err; echo "bench tuning rc=$?"; cut -c1-260 gpurun_out/r2c8_bench_tuning.json; tail -2 gpurun_out/r2c8_bench_tuning.err

#!/bin/bash
# round-2 GPU call: full gpu test-suite with the new defaults + two-tile forward, pipe-rate microbench, attention
# variant timing, short bench
mkdir -p gpurun_out
timeout 60 tools/micro/pipe_rates > gpurun_out/r2_pipe_rates.log 2>&1; echo "pipe_rates rc=$?"; cat gpurun_out/r2_pipe_rates.log | grep -E "16 |status"
timeout 240 python tools/attn_bench.py > gpurun_out/r2_attn_bench.log 2>&1; echo "attn_bench rc=$?"; cut -c1-420 gpurun_out/r2_attn_bench.log | tail -12
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -p no:cacheprovider > gpurun_out/r2_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -15 gpurun_out/r2_gputests.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2_bench_short.json; tail -3 gpurun_out/r2_bench_short.err

#!/bin/bash
# validation of the elect_one() issue path: kernel parity tests, then a short bench, then kernel micro-timings
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py -x -q > gpurun_out/el_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/el_tests.log
tail -3 gpurun_out/el_tests.log
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-micro --no-e2e > gpurun_out/el_bench.json 2> gpurun_out/el_bench.err; echo "bench rc=$?"
cat gpurun_out/el_bench.json
timeout 60 python tools/bench_kernels.py attn gemm > gpurun_out/el_kernels.log 2>&1; tail -32 gpurun_out/el_kernels.log
PP_MODES=1,4 timeout 60 python tools/pp_sweep.py > gpurun_out/pp_sweep2.log 2>&1; tail -3 gpurun_out/pp_sweep2.log | cut -c1-400

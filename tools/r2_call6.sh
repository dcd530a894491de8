#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_variants_gpu.py tests/test_gemm_gpu.py -m gpu -x -q --timeout 120 -p no:cacheprovider > gpurun_out/r2c6_variants.log 2>&1; echo "variants+gemm rc=$?"; tail -3 gpurun_out/r2c6_variants.log
timeout 240 python tools/attn_bench.py fwd > gpurun_out/r2c6_attn_bench.log 2>&1; echo "attn_bench rc=$?"
python - <<'PY'
import json
for r in json.load(open("gpurun_out/attn_bench.json")):
    print(r["shape"], {k: (v["ms"], v["frac"], round(v["err_vs_legacy"], 5)) for k, v in r.items() if k.startswith("fwd")})
PY
timeout 120 ncu --set full --import-source on --clock-control none -k regex:attn_fwd2 -c 1 -o gpurun_out/r2c6_fwd2 -f python tools/prof_attn.py > gpurun_out/r2c6_ncu.log 2>&1; echo "ncu rc=$?"
timeout 200 python tools/gemm_probe.py > gpurun_out/r2c6_gemm_probe.log 2>&1; echo "gemm_probe rc=$?"; cat gpurun_out/r2c6_gemm_probe.log | cut -c1-260
timeout 900 python -m pytest tests/test_tuning_gpu.py tests/test_dropin_gpu.py tests/test_pipeline_gpu.py -m gpu -q -s --timeout 300 -p no:cacheprovider > gpurun_out/r2c6_tuning.log 2>&1; echo "tuning/dropin/pipeline rc=$?"; grep -E "^\[|passed|failed|^FAILED|^E  " gpurun_out/r2c6_tuning.log | cut -c1-300 | head -40
timeout 600 python -m pytest tests/test_e2e_gpu.py tests/test_kernels_gpu.py -m gpu -q -s --timeout 400 -p no:cacheprovider > gpurun_out/r2c6_e2e.log 2>&1; echo "e2e+kernels rc=$?"; grep -E "^\[|^\.\[|passed|failed|^FAILED|^E  " gpurun_out/r2c6_e2e.log | cut -c1-400 | head -30
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/r2c6_bench.json 2> gpurun_out/r2c6_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2c6_bench.json; tail -3 gpurun_out/r2c6_bench.err

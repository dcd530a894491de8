#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_variants_gpu.py -m gpu -x -q --timeout 120 -p no:cacheprovider -k "forward_variants" > gpurun_out/r2c5_variants.log 2>&1; echo "variants rc=$?"; tail -3 gpurun_out/r2c5_variants.log
timeout 240 python tools/attn_bench.py fwd > gpurun_out/r2c5_attn_bench.log 2>&1; echo "attn_bench rc=$?"
python - <<'PY'
import json
for r in json.load(open("gpurun_out/attn_bench.json"))[:2]:
    print(r["shape"], {k: (v["ms"], v["frac"], round(v["err_vs_legacy"], 5)) for k, v in r.items() if k.startswith("fwd")})
PY
E4T_ATTN_FWD2=p0 timeout 120 ncu --set full --import-source on --clock-control none -k regex:attn_fwd2 -c 1 -o gpurun_out/r2c5_fwd2_p0 -f python tools/prof_attn.py > gpurun_out/r2c5_ncu_p0.log 2>&1; echo "ncu rc=$?"
timeout 600 python -m pytest tests/test_tuning_gpu.py tests/test_dropin_gpu.py -m gpu -q -s --timeout 300 -p no:cacheprovider > gpurun_out/r2c5_tuning.log 2>&1; echo "tuning tests rc=$?"; grep -E "^\[|passed|failed|Error|FAILED|assert" gpurun_out/r2c5_tuning.log | cut -c1-300 | head -40
timeout 200 python tools/gemm_probe.py > gpurun_out/r2c5_gemm_probe.log 2>&1; echo "gemm_probe rc=$?"; cat gpurun_out/r2c5_gemm_probe.log | cut -c1-260
timeout 500 python -m pytest tests/test_e2e_gpu.py tests/test_kernels_gpu.py -m gpu -x -q --timeout 400 -p no:cacheprovider > gpurun_out/r2c5_e2e.log 2>&1; echo "e2e+kernels rc=$?"; tail -4 gpurun_out/r2c5_e2e.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2c5_bench.json; tail -3 gpurun_out/r2c5_bench.err
timeout 300 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -s --timeout 200 -p no:cacheprovider > gpurun_out/r2c5_pipeline.log 2>&1; echo "pipeline rc=$?"; grep -E "^\[|passed|failed|Error" gpurun_out/r2c5_pipeline.log | cut -c1-300
timeout 120 ncu --set full --import-source on --clock-control none -k regex:gn_ -s 8 -c 8 -o gpurun_out/r2c5_gn -f python tools/prof_gn.py > gpurun_out/r2c5_ncu_gn.log 2>&1; echo "ncu gn rc=$?"

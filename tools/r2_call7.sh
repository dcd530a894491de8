#!/bin/bash
mkdir -p gpurun_out
for p in 0 1; do
  E4T_GEMM_POLL=$p timeout 200 python tools/gemm_probe.py > gpurun_out/r2c7_gemm_probe_poll$p.log 2>&1; echo "gemm_probe poll=$p rc=$?"; cut -c1-200 gpurun_out/r2c7_gemm_probe_poll$p.log
done
timeout 200 python tools/sweep_r2.py norm > gpurun_out/r2c7_norm.log 2>&1; echo "norm sweep rc=$?"; grep -E "^\[norm\]" gpurun_out/r2c7_norm.log | cut -c1-220
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_variants_gpu.py -m gpu -x -q --timeout 200 -p no:cacheprovider > gpurun_out/r2c7_kernels.log 2>&1; echo "kernel tests rc=$?"; tail -3 gpurun_out/r2c7_kernels.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c7_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches_per_step", "loss")})
    print("e2e", d["e2e"]); print("clocks", d["clocks"])
    r = d["roofline"]; print("roofline", {k: r[k] for k in ("achieved", "frac", "traffic")})
    a = r["aggregate_wo_attention"]; print("aggregate", {k: a[k] for k in ("gflop_per_image_per_step", "ms_per_step", "achieved", "frac")}); print(a["by_level"])
    print("kernels", {k: {kk: round(vv, 3) for kk, vv in v.items() if isinstance(vv, float)} for k, v in d["kernels"].items()})
    print("cpu", d["cpu_baseline"])
except Exception as ex:
    print("parse failed", ex); print(open("gpurun_out/r2c7_bench.err").read()[-1500:])
PY
timeout 400 python bench.py --impl torch_stock --steps 3 --warmup 2 > gpurun_out/r2c7_torch_stock.json 2> gpurun_out/r2c7_torch_stock.err; echo "torch_stock rc=$?"; cut -c1-400 gpurun_out/r2c7_torch_stock.json; tail -3 gpurun_out/r2c7_torch_stock.err | cut -c1-300
timeout 120 ncu --set full --import-source on --clock-control none -k regex:attn_fwd2 -c 1 -o gpurun_out/r2c7_fwd2 -f python tools/prof_attn.py > gpurun_out/r2c7_ncu.log 2>&1; echo "ncu fwd2 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv python bench.py --profile-one-step --warmup 1 --no-cpu-baseline > gpurun_out/r2c7_launches.log 2>&1; echo "launch list rc=$?"

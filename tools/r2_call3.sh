#!/bin/bash
# round-2 GPU call 3: fwd2 variants (exp-phase token, fp16 P) parity + timing, ncu of the shipped fwd2, measured e2e errors
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_variants_gpu.py -m gpu -x -q --timeout 120 -p no:cacheprovider -k "forward_variants" > gpurun_out/r2c3_variants.log 2>&1; echo "variants rc=$?"; tail -6 gpurun_out/r2c3_variants.log
timeout 240 python tools/attn_bench.py fwd > gpurun_out/r2c3_attn_bench.log 2>&1; echo "attn_bench rc=$?"
python - <<'PY'
import json
for r in json.load(open("gpurun_out/attn_bench.json")):
    print(r["shape"], {k: (v["ms"], v["frac"], round(v["err_vs_legacy"], 5)) for k, v in r.items() if k.startswith("fwd")})
PY
for v in "p0" "b"; do
  E4T_ATTN_FWD2=$v timeout 120 ncu --set full --import-source on --clock-control none -k regex:attn_fwd2 -c 1 \
      -o gpurun_out/r2c3_fwd2_$v -f python tools/prof_attn.py > gpurun_out/r2c3_ncu_$v.log 2>&1
  echo "ncu $v rc=$?"
done

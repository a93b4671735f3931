#!/bin/bash
# round 2, GPU session 18: weight-stationary GEMM slice width A/B (whole step)
mkdir -p gpurun_out
S=gpurun_out/r2s18
for v in 0 16777216 33554432 50331648; do
  DTLR_DEBUG_FLAGS=$v timeout 900 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench_$v.json 2> ${S}_bench_$v.err; echo "bench flags $v rc $?"
done
python - <<'PY'
import json
for t in ("0", "16777216", "33554432", "50331648"):
    try:
        d = json.load(open("gpurun_out/r2s18_bench_%s.json" % t))
        print(t, {k: d[k] for k in ("value", "ms_per_step")}, "gemm frac", d["roofline_gemm"]["frac"], d["roofline_gemm"]["hbm_view"]["frac"])
    except Exception as e:
        print(t, "failed", e)
PY

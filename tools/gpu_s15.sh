#!/bin/bash
# round 2, GPU session 15: A/B of the LayerNorm-fused K=256 GEMM (DTLR_LN_FUSE_MIN_K=256), fresh launch list
mkdir -p gpurun_out
S=gpurun_out/r2s15
for v in base lnfuse256; do
  if [ $v = lnfuse256 ]; then export DTLR_LN_FUSE_MIN_K=256; fi
  timeout 900 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench_$v.json 2> ${S}_bench_$v.err; echo "bench $v rc $?"
done
unset DTLR_LN_FUSE_MIN_K
python - <<'PY'
import json
for t in ("base", "lnfuse256"):
    try:
        d = json.load(open("gpurun_out/r2s15_bench_%s.json" % t))
        print(t, {k: d[k] for k in ("value", "ms_per_step", "dtype", "gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e_u8"]["value"])
    except Exception as e:
        print(t, "failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${S}_launches.csv python tools/profile_step.py 2 > ${S}_ll.log 2>&1; echo "launch list rc $?"

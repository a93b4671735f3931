#!/bin/bash
# round 2, GPU session 17: ncu --set full (with source) of the K=256 projection kernels, add_layernorm, the box head; bench with MSDA unroll 4
mkdir -p gpurun_out
S=gpurun_out/r2s17
timeout 600 ncu --set full --clock-control none --import-source on -s 14 -c 7 -f -o ${S}_small python tools/profile_small.py > ${S}_ncu_small.log 2>&1; echo "ncu small rc $?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s17_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], "msda", d["roofline_msda"]["us_per_launch"], d["roofline_msda"]["frac"])
PY

#!/bin/bash
# round 2, GPU session 9: attention auto-partition, shape-constant caches; suite, bench, fresh launch list
mkdir -p gpurun_out
S=gpurun_out/r2s9
timeout 600 python -m pytest tests -m gpu -q > ${S}_tests.log 2>&1; echo "tests rc $?"; tail -4 ${S}_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s9_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "dtype", "gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e_u8"]["value"], "ffn", d["roofline"]["us_per_launch"], d["roofline"]["frac"], "msda", d["roofline_msda"]["us_per_launch"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${S}_launches.csv python tools/profile_step.py 2 > ${S}_ll.log 2>&1; echo "launch list rc $?"

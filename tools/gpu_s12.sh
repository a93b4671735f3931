#!/bin/bash
# round 2, GPU session 12: attention kernel back to the spill-free 16-warp build; tests, sweep, bench
mkdir -p gpurun_out
S=gpurun_out/r2s12
timeout 300 python -m pytest tests/test_gpu_attention.py tests/test_gpu_engine.py -m gpu -q > ${S}_unit.log 2>&1; echo "attention+engine tests rc $?"; tail -3 ${S}_unit.log
timeout 300 python tools/bench_attn.py > ${S}_attn.log 2>&1; cat ${S}_attn.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s12_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "dtype", "gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e_u8"]["value"], "ffn", d["roofline"]["us_per_launch"], "msda", d["roofline_msda"]["us_per_launch"])
PY

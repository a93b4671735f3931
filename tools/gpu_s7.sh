#!/bin/bash
# round 2, GPU session 7: flash-attention work partition sweep + parity
mkdir -p gpurun_out
S=gpurun_out/r2s7
timeout 300 python -m pytest tests/test_gpu_attention.py -m gpu -q > ${S}_attn_tests.log 2>&1; echo "attention tests rc $?"; tail -3 ${S}_attn_tests.log
timeout 600 python tools/bench_attn.py > ${S}_attn.log 2>&1; cat ${S}_attn.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s7_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "dtype", "gpu_launches")}, "e2e", d["e2e"]["value"], "ffn", d["roofline"]["us_per_launch"], d["roofline"]["frac"])
PY

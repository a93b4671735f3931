#!/bin/bash
# round 2, GPU session 14: conv producer without per-k-step divisions; conv tests, timing table, suite, bench
mkdir -p gpurun_out
S=gpurun_out/r2s14
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_gemm.py -m gpu -q > ${S}_unit.log 2>&1; echo "conv+gemm tests rc $?"; tail -3 ${S}_unit.log
timeout 300 python tools/profile_conv.py time > ${S}_conv_times.log 2>&1; cat ${S}_conv_times.log
timeout 600 python -m pytest tests -m gpu -q > ${S}_tests.log 2>&1; echo "tests rc $?"; tail -3 ${S}_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s14_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "dtype", "gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e_u8"]["value"], "ffn", d["roofline"]["us_per_launch"], "msda", d["roofline_msda"]["us_per_launch"])
PY

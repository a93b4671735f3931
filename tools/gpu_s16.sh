#!/bin/bash
# round 2, GPU session 16: MSDA phase-2 unroll A/B, parity, bench
mkdir -p gpurun_out
S=gpurun_out/r2s16
timeout 300 python -m pytest tests/test_gpu_msda.py tests/test_gpu_msda_vs_ref_cuda.py -m gpu -q > ${S}_msda_tests.log 2>&1; echo "msda tests rc $?"; tail -3 ${S}_msda_tests.log
DTLR_DEBUG_FLAGS=8388608 timeout 300 python -m pytest tests/test_gpu_msda.py -m gpu -q > ${S}_msda_tests_qu4.log 2>&1; echo "msda tests qu4 rc $?"; tail -2 ${S}_msda_tests_qu4.log
timeout 300 python tools/bench_msda.py > ${S}_msda.log 2>&1; head -12 ${S}_msda.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"
DTLR_DEBUG_FLAGS=4194304 timeout 900 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench_qu1.json 2> ${S}_bench_qu1.err
DTLR_DEBUG_FLAGS=8388608 timeout 900 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench_qu4.json 2> ${S}_bench_qu4.err
python - <<'PY'
import json
for t in ("bench", "bench_qu1", "bench_qu4"):
    try:
        d = json.load(open("gpurun_out/r2s16_%s.json" % t))
        print(t, {k: d[k] for k in ("value", "ms_per_step")}, "msda", d["roofline_msda"]["us_per_launch"], d["roofline_msda"]["frac"])
    except Exception as e:
        print(t, "failed", e)
PY

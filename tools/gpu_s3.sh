#!/bin/bash
# round 2, GPU session 3: fp16 flavour (libdtlr_b200_f16.so): full suite under bf16 (default) and with the 16-bit tests aliased to fp16,
# bench in both 16-bit modes
mkdir -p gpurun_out
S=gpurun_out/r2s3
timeout 600 python -m pytest tests -m gpu -q > ${S}_tests.log 2>&1; echo "tests rc $?"; tail -8 ${S}_tests.log
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -s -k "throughput or hwdb" > ${S}_modes.log 2>&1; grep -E "B=64|HWDB|float16|bfloat16" ${S}_modes.log | head -20
DTLR_TEST_HALF=f16 timeout 600 python -m pytest tests -m gpu -q -x > ${S}_tests_f16.log 2>&1; echo "f16-aliased tests rc $?"; tail -8 ${S}_tests_f16.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench_f16.json 2> ${S}_bench_f16.err; echo "bench f16 rc $?"
timeout 900 python bench.py --steps 10 --warmup 3 --dtype bf16 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench_bf16.json 2> ${S}_bench_bf16.err; echo "bench bf16 rc $?"
python - <<'PY'
import json
for t in ("f16", "bf16"):
    try:
        d = json.load(open("gpurun_out/r2s3_bench_%s.json" % t))
        print(t, {k: d[k] for k in ("value", "ms_per_step", "dtype", "gpu_launches")}, "e2e", d["e2e"]["value"], "ffn", d["roofline"]["us_per_launch"], d["roofline"]["frac"], "msda", d["roofline_msda"]["us_per_launch"])
    except Exception as e:
        print(t, "failed", e)
PY
tail -3 ${S}_bench_f16.err
python __graft_entry__.py smoke 2>&1 | tail -8

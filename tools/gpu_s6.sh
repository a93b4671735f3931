#!/bin/bash
# round 2, GPU session 6: FFN wave-quantisation tail split (PART variant) -- unit tests, suite, A/B bench
mkdir -p gpurun_out
S=gpurun_out/r2s6
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -k "ffn or mlp_head" > ${S}_ffn.log 2>&1; echo "ffn tests rc $?"; tail -5 ${S}_ffn.log
DTLR_TEST_HALF=f16 timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -k "ffn or mlp_head" > ${S}_ffn_f16.log 2>&1; echo "ffn tests f16 rc $?"; tail -3 ${S}_ffn_f16.log
timeout 600 python -m pytest tests -m gpu -q > ${S}_tests.log 2>&1; echo "tests rc $?"; tail -4 ${S}_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"
DTLR_FFN_SPLIT_TAIL=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench_nosplit.json 2> ${S}_bench_nosplit.err
python - <<'PY'
import json
for t in ("bench", "bench_nosplit"):
    try:
        d = json.load(open("gpurun_out/r2s6_%s.json" % t))
        print(t, {k: d[k] for k in ("value", "ms_per_step", "dtype", "gpu_launches")}, "e2e", d["e2e"]["value"], "ffn", d["roofline"]["us_per_launch"], d["roofline"]["frac"])
    except Exception as e:
        print(t, "failed", e)
PY
python tools/profile_ffn.py > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ffn -c 12 python tools/profile_ffn.py 2>&1 | grep -E "ffn_|gpu__time" | head -30

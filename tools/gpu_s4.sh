#!/bin/bash
# round 2, GPU session 4: fused MLP head (dtlr_mlp_head) + full suite in both 16-bit flavours + bench
mkdir -p gpurun_out
S=gpurun_out/r2s4
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -k mlp_head > ${S}_head.log 2>&1; echo "mlp_head rc $?"; tail -5 ${S}_head.log
timeout 600 python -m pytest tests -m gpu -q > ${S}_tests.log 2>&1; echo "tests rc $?"; tail -6 ${S}_tests.log
DTLR_TEST_HALF=f16 timeout 600 python -m pytest tests -m gpu -q > ${S}_tests_f16.log 2>&1; echo "f16-aliased tests rc $?"; tail -12 ${S}_tests_f16.log
timeout 900 python bench.py --steps 10 --warmup 3 > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"
DTLR_MLP_HEAD_FUSED=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench_nohead.json 2> ${S}_bench_nohead.err
python - <<'PY'
import json
for t in ("bench", "bench_nohead"):
    try:
        d = json.load(open("gpurun_out/r2s4_%s.json" % t))
        print(t, {k: d[k] for k in ("value", "ms_per_step", "dtype", "gpu_launches")}, "e2e", d["e2e"]["value"], "ffn", d["roofline"]["us_per_launch"], d["roofline"]["frac"], "train", d.get("train_step"))
    except Exception as e:
        print(t, "failed", e)
PY

#!/bin/bash
# round 2, GPU session 11: strided implicit-GEMM convs (TMA traversal stride), attention tail-block skip; tests, A/B bench
mkdir -p gpurun_out
S=gpurun_out/r2s11
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_attention.py -m gpu -q > ${S}_unit.log 2>&1; echo "conv+attention tests rc $?"; tail -4 ${S}_unit.log
timeout 600 python -m pytest tests -m gpu -q > ${S}_tests.log 2>&1; echo "tests rc $?"; tail -3 ${S}_tests.log
DTLR_TEST_HALF=f16 timeout 600 python -m pytest tests -m gpu -q > ${S}_tests_f16.log 2>&1; echo "f16-aliased tests rc $?"; tail -3 ${S}_tests_f16.log
timeout 300 python tools/bench_attn.py 2>&1 | head -3
timeout 900 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"
DTLR_CONV_STRIDED_IMPLICIT=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench_im2col.json 2> ${S}_bench_im2col.err
python - <<'PY'
import json
for t in ("bench", "bench_im2col"):
    try:
        d = json.load(open("gpurun_out/r2s11_%s.json" % t))
        print(t, {k: d[k] for k in ("value", "ms_per_step", "dtype", "gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e_u8"]["value"], "ffn", d["roofline"]["us_per_launch"], "msda", d["roofline_msda"]["us_per_launch"])
    except Exception as e:
        print(t, "failed", e)
PY

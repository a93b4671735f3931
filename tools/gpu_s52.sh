#!/bin/bash
# round 2, GPU session 52: tcgen05 attention starts on the first 256 keys (per-box K / V barriers)
mkdir -p gpurun_out
S=gpurun_out/r2s52
timeout 200 python -m pytest tests/test_gpu_attention.py -q -m gpu -x > ${S}_tests.txt 2>&1; rc=$?; echo "attention tests rc $rc"; tail -3 ${S}_tests.txt | cut -c1-300
DTLR_TEST_HALF=f16 timeout 200 python -m pytest tests/test_gpu_attention.py tests/test_gpu_engine.py -q -m gpu -x > ${S}_tests_f16.txt 2>&1; echo "f16 attention+engine rc $?"; tail -3 ${S}_tests_f16.txt | cut -c1-300
timeout 150 python tools/attn_timeline.py 2>&1 | head -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"; cut -c1-200 ${S}_bench.json

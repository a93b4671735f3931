#!/bin/bash
# round 2, GPU session 49: tcgen05 attention as the default -- full GPU suite (both flavour aliases for the engine), bench A/B vs the flash kernel
mkdir -p gpurun_out
S=gpurun_out/r2s49
timeout 900 python -m pytest tests -q -m gpu -x > ${S}_suite.txt 2>&1; echo "suite rc $?"; tail -4 ${S}_suite.txt | cut -c1-300
DTLR_TEST_HALF=f16 timeout 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_attention.py -q -m gpu -x > ${S}_suite_f16.txt 2>&1; echo "f16 engine+attention rc $?"; tail -3 ${S}_suite_f16.txt | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"; cut -c1-200 ${S}_bench.json
DTLR_ATTN=hmma timeout 600 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench_hmma.json 2> ${S}_bench_hmma.err; echo "bench (flash attention) rc $?"; cut -c1-200 ${S}_bench_hmma.json

#!/bin/bash
# round 2, GPU session 53: attention edge shapes (dead warps, 1-key / 16-key last chunks, 1..8 tiles), full suite
mkdir -p gpurun_out
S=gpurun_out/r2s53
timeout 900 python -m pytest tests -q -m gpu -x > ${S}_suite.txt 2>&1; echo "suite rc $?"; tail -4 ${S}_suite.txt | cut -c1-300
DTLR_TEST_HALF=f16 timeout 300 python -m pytest tests/test_gpu_attention.py tests/test_gpu_gemm.py -q -m gpu -x > ${S}_suite_f16.txt 2>&1; echo "f16 attention+gemm rc $?"; tail -3 ${S}_suite_f16.txt | cut -c1-300

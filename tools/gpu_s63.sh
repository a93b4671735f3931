#!/bin/bash
# round 2, GPU session 63: column-tiles-fastest order for the split-precision products (A shared through L2), A/B against the old order
mkdir -p gpurun_out
S=gpurun_out/r2s63
timeout 200 python -m pytest tests/test_gpu_split.py tests/test_gpu_gemm.py -x -q > ${S}_tests.txt 2>&1; echo "split + gemm tests rc $?"; tail -2 ${S}_tests.txt | cut -c1-300
timeout 200 python tools/bench_split.py shapes > ${S}_shapes.txt 2>&1; echo "bench_split rc $?"; grep -a "split mode\|contractions\| us " ${S}_shapes.txt | head -8 | cut -c1-180
DTLR_DEBUG_FLAGS=134217728 timeout 100 python tools/bench_split.py > ${S}_rowfast.txt 2>&1; tail -1 ${S}_rowfast.txt

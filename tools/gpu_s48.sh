#!/bin/bash
# round 2, GPU session 48: tcgen05 attention -- issuers poll the four tile pipelines, staggered start
mkdir -p gpurun_out
S=gpurun_out/r2s48
timeout 200 python -m pytest tests/test_gpu_attention.py -q -m gpu -x > ${S}_tests.txt 2>&1; rc=$?; echo "attention tests rc $rc"; tail -6 ${S}_tests.txt | cut -c1-300
DTLR_TEST_HALF=f16 timeout 200 python -m pytest tests/test_gpu_attention.py -q -m gpu -x > ${S}_tests_f16.txt 2>&1; echo "attention tests f16 rc $?"; tail -3 ${S}_tests_f16.txt | cut -c1-300
timeout 150 python tools/attn_timeline.py > ${S}_attn_timeline.txt 2>&1; echo rc $?; cut -c1-170 ${S}_attn_timeline.txt | head -64

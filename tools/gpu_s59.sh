#!/bin/bash
# round 2, GPU session 59: split-precision products with hi / lo tiles loaded once (DTLR_SPLIT16 in: e.split3), A/B against the plain 3K walk
mkdir -p gpurun_out
S=gpurun_out/r2s59
timeout 200 python -m pytest tests/test_gpu_split.py -x -q > ${S}_split_tests.txt 2>&1; echo "split kernel tests rc $?"; tail -4 ${S}_split_tests.txt | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_engine.py -x -q -s -k "split" > ${S}_engine_tests.txt 2>&1; echo "engine split tests rc $?"; grep -a "split\|passed\|failed\|Error\|assert" ${S}_engine_tests.txt | cut -c1-400 | tail -16
timeout 200 python tools/bench_split.py table > ${S}_split_table.txt 2>&1; echo "bench_split rc $?"; grep -a "split mode\|eager step\| us " ${S}_split_table.txt | head -9 | cut -c1-180
DTLR_SPLIT3_LOADS=0 timeout 100 python tools/bench_split.py > ${S}_split_plainwalk.txt 2>&1; tail -1 ${S}_split_plainwalk.txt

#!/bin/bash
# round 2, GPU session 58: split-precision mode with split-output epilogues (DTLR_SPLIT16), shared block-input split, direct 16-bit q / k / v
mkdir -p gpurun_out
S=gpurun_out/r2s58
timeout 200 python -m pytest tests/test_gpu_split.py -q > ${S}_split_tests.txt 2>&1; echo "split kernel tests rc $?"; tail -4 ${S}_split_tests.txt | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_engine.py -x -q -s -k "split" > ${S}_engine_tests.txt 2>&1; echo "engine split tests rc $?"; grep -a "split\|passed\|failed\|Error\|assert" ${S}_engine_tests.txt | cut -c1-400 | tail -25
timeout 200 python tools/bench_split.py table > ${S}_split_table.txt 2>&1; echo "bench_split rc $?"; grep -a "split mode\|eager step\| us " ${S}_split_table.txt | head -22 | cut -c1-180
DTLR_SPLIT_OUT_FUSED=0 timeout 100 python tools/bench_split.py > ${S}_split_unfused.txt 2>&1; tail -1 ${S}_split_unfused.txt

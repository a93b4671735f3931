#!/bin/bash
# round 2, GPU session 57: first run of the split-precision mode (tensor-core parity mode): kernel tests, engine tests vs fixture / oracle,
# timing at the bench shape with a kernel table, implicit-conv A/B
mkdir -p gpurun_out
S=gpurun_out/r2s57
timeout 200 python -m pytest tests/test_gpu_split.py -x -q -s > ${S}_split_tests.txt 2>&1; echo "split kernel tests rc $?"; tail -3 ${S}_split_tests.txt | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_engine.py -x -q -s -k "split or bench_shape_fp32" > ${S}_engine_tests.txt 2>&1; echo "engine split tests rc $?"; grep -a "split\|passed\|failed\|Error\|assert" ${S}_engine_tests.txt | cut -c1-400 | tail -25
timeout 200 python tools/bench_split.py table > ${S}_split_table.txt 2>&1; echo "bench_split rc $?"; head -34 ${S}_split_table.txt | cut -c1-200
DTLR_SPLIT_CONV_IMPLICIT=0 timeout 100 python tools/bench_split.py > ${S}_split_im2col.txt 2>&1; tail -1 ${S}_split_im2col.txt

#!/bin/bash
# round 2, GPU session 24: native front (ResNet layer2-4 + input_proj backward kernels), engine parity, timing
mkdir -p gpurun_out
S=gpurun_out/r2s24
timeout 900 python -m pytest tests/test_gpu_train_kernels.py -q -m gpu -s -k "groupnorm or col2im or dual or front" > ${S}_front.txt 2>&1; echo "front tests rc $?"; grep "native front\|passed\|failed\|Error\|error" ${S}_front.txt | cut -c1-300 | head -20
timeout 900 python -m pytest tests/test_gpu_train_engine.py -q -m gpu -s > ${S}_engine.txt 2>&1; echo "engine tests rc $?"; grep "worst\|bf16 loss\|passed\|failed\|Error" ${S}_engine.txt | cut -c1-300
timeout 600 python tools/bench_train_native.py 32 bf16 > ${S}_train.txt 2>&1; echo "timing rc $?"; grep "variant\|Error" ${S}_train.txt
DTLR_DEBUG_FLAGS=64 DTLR_TRAIN_PROFILE=1 timeout 900 python tools/bench_train_native.py 32 bf16 > ${S}_train_nopdl.txt 2>&1; echo "profile rc $?"; grep -A40 "GPU kernel time" ${S}_train_nopdl.txt | cut -c1-170

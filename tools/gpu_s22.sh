#!/bin/bash
# round 2, GPU session 22: flash self-attention forward (mask, lse) + backward kernels, colsum / LayerNorm-backward reduction fix
mkdir -p gpurun_out
S=gpurun_out/r2s22
timeout 900 python -m pytest tests/test_gpu_train_kernels.py -q -m gpu > ${S}_kernels.txt 2>&1; echo "kernel tests rc $?"; tail -25 ${S}_kernels.txt | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_train_engine.py -q -m gpu -s > ${S}_engine.txt 2>&1; echo "engine tests rc $?"; grep "worst\|bf16 loss\|passed\|failed\|Error" ${S}_engine.txt | cut -c1-300
timeout 900 python tools/bench_train_native.py 32 bf16 > ${S}_train.txt 2>&1; echo "timing rc $?"; grep variant ${S}_train.txt
DTLR_DEBUG_FLAGS=64 DTLR_TRAIN_PROFILE=1 timeout 900 python tools/bench_train_native.py 32 bf16 > ${S}_train_nopdl.txt 2>&1; echo "profile rc $?"; grep -A36 "GPU kernel time" ${S}_train_nopdl.txt | cut -c1-170

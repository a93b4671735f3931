#!/bin/bash
# round 2, GPU session 20: native fine-tune step bring-up -- wgrad descriptor probe, backward-kernel tests, engine parity, timing
mkdir -p gpurun_out
S=gpurun_out/r2s20
timeout 300 python tools/probe_wgrad.py > ${S}_probe.txt 2>&1; echo "probe rc $?"; cat ${S}_probe.txt | tail -8
timeout 900 python -m pytest tests/test_gpu_train_kernels.py -q -x -m gpu > ${S}_kernels.txt 2>&1; echo "kernel tests rc $?"; tail -15 ${S}_kernels.txt
timeout 900 python -m pytest tests/test_gpu_train_engine.py -q -m gpu -s > ${S}_engine.txt 2>&1; echo "engine tests rc $?"; tail -30 ${S}_engine.txt
DTLR_TRAIN_PROFILE=1 timeout 900 python tools/bench_train_native.py 32 bf16,torch > ${S}_train.txt 2>&1; echo "timing rc $?"; head -60 ${S}_train.txt | cut -c1-180

#!/bin/bash
# round 2, GPU session 27: fused ReLU-backward dgrad epilogue, ragged-batch parity, ncu launch list + full captures of the new backward kernels
mkdir -p gpurun_out
S=gpurun_out/r2s27
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_engine.py tests/test_gpu_gemm.py tests/test_gpu_conv.py -q -m gpu -s > ${S}_tests.txt 2>&1; echo "tests rc $?"; grep "worst\|bf16 loss\|ragged\|passed\|failed\|Error\|native front" ${S}_tests.txt | cut -c1-300
timeout 600 python tools/bench_train_native.py 32 bf16 > ${S}_train.txt 2>&1; echo "timing: $(grep variant ${S}_train.txt)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${S}_launches.csv python tools/profile_train_kernels.py > /dev/null 2>&1; echo "ncu list rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tcgen05|sa_bwd_dkv|sa_bwd_dq|sa_fwd" -s 12 -c 6 -o ${S}_train_kernels python tools/profile_train_kernels.py > /dev/null 2>&1; echo "ncu full rc $?"
ls -la gpurun_out | tail -5

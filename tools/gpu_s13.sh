#!/bin/bash
# round 2, GPU session 13: conv timing table + ncu of the layer1 / layer4 conv kernels
mkdir -p gpurun_out
S=gpurun_out/r2s13
timeout 300 python tools/profile_conv.py time > ${S}_conv_times.log 2>&1; cat ${S}_conv_times.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05_kernel -s 16 -c 4 -f -o ${S}_conv python tools/profile_conv.py > ${S}_ncu_conv.log 2>&1; echo "ncu conv rc $?"

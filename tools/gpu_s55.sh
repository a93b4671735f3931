#!/bin/bash
# round 2, GPU session 55: fresh ncu --set full captures at the final HEAD: MSDA forward, the K = 256 projection kernels, the 3x3 convs; conv timing table
mkdir -p gpurun_out
S=gpurun_out/r2s55
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 4 -c 1 -f -o ${S}_msda python tools/profile_msda.py > ${S}_ncu_msda.log 2>&1; echo "ncu msda rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_ws|add_layernorm|ffn_ln_tcgen05" -s 6 -c 6 -f -o ${S}_small python tools/profile_small.py > ${S}_ncu_small.log 2>&1; echo "ncu small rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -c 8 -f -o ${S}_conv python tools/profile_conv.py > ${S}_ncu_conv.log 2>&1; echo "ncu conv rc $?"
timeout 200 python tools/profile_conv.py time > ${S}_conv_times.txt 2>&1; cat ${S}_conv_times.txt | cut -c1-200

#!/bin/bash
# round 2, GPU session 30: where the stream-K FFN kernels spend their time (parts switched off by probe flags), both CTA modes
mkdir -p gpurun_out
S=gpurun_out/r2s30
FFN_PROBES=1 timeout 600 python tools/bench_ffn.py 58368 > ${S}_ffn_probes.txt 2>&1; cat ${S}_ffn_probes.txt
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv

#!/bin/bash
# round 2, GPU session 33: timeline of the CTA-pair stream-K FFN kernel
mkdir -p gpurun_out
timeout 120 python tools/ffn_timeline.py > gpurun_out/r2s33_ffn_timeline.txt 2>&1; echo rc $?; head -75 gpurun_out/r2s33_ffn_timeline.txt | cut -c1-210

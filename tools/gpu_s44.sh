#!/bin/bash
# round 2, GPU session 44: timeline of the single-pass tcgen05 attention kernel
mkdir -p gpurun_out
timeout 150 python tools/attn_timeline.py > gpurun_out/r2s44_attn_timeline.txt 2>&1; echo rc $?; cut -c1-170 gpurun_out/r2s44_attn_timeline.txt | head -80

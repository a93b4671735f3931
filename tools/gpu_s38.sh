#!/bin/bash
# round 2, GPU session 38: timeline of the weight-stationary GEMM kernel
mkdir -p gpurun_out
timeout 100 python tools/ws_timeline.py 58368 256 256 0 > gpurun_out/r2s38_ws_timeline.txt 2>&1; echo rc $?
timeout 100 python tools/ws_timeline.py 58368 256 256 1 >> gpurun_out/r2s38_ws_timeline.txt 2>&1; echo rc $?
timeout 100 python tools/ws_timeline.py 58368 384 256 0 >> gpurun_out/r2s38_ws_timeline.txt 2>&1; echo rc $?
cat gpurun_out/r2s38_ws_timeline.txt | cut -c1-200

#!/bin/bash
# round 2, GPU session 43: weight-stationary GEMM epilogue with TMEM loads one 32-column piece ahead -- parity, timing per shape, step
mkdir -p gpurun_out
S=gpurun_out/r2s43
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_conv.py -q -m gpu -x > ${S}_tests.txt 2>&1; rc=$?; echo "gemm+conv tests rc $rc"; tail -12 ${S}_tests.txt | cut -c1-300
if [ $rc -eq 0 ]; then
  for shape in "58368 256 256 0" "58368 256 256 1" "58368 384 256 0" "57600 512 256 0" "57600 166 256 0" "163840 256 64 0"; do
    WS_TIME_ONLY=1 timeout 60 python tools/ws_timeline.py $shape 0 2>&1 | tail -1
  done > ${S}_ws_times.txt; cat ${S}_ws_times.txt
  timeout 100 python tools/ws_timeline.py 58368 256 256 0 > ${S}_ws_timeline.txt 2>&1; sed -n 2,20p ${S}_ws_timeline.txt | cut -c1-160
  timeout 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -x > ${S}_engine_tests.txt 2>&1; echo "engine tests rc $?"; tail -3 ${S}_engine_tests.txt | cut -c1-300
  timeout 600 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"; cut -c1-200 ${S}_bench.json
fi

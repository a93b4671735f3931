#!/bin/bash
# round 2, GPU session 37: box head on the pipelined CTA-pair kernel -- parity, timing, ncu of the FFN kernel
mkdir -p gpurun_out
S=gpurun_out/r2s37
timeout 120 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "mlp_head" -x > ${S}_tests.txt 2>&1; rc=$?; echo "tests rc $rc"; tail -12 ${S}_tests.txt | cut -c1-300
if [ $rc -eq 0 ]; then
  timeout 120 python tools/bench_head.py > ${S}_head_bench.txt 2>&1; cat ${S}_head_bench.txt
  timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_train_engine.py -q -m gpu -x > ${S}_engine_tests.txt 2>&1; echo "engine tests rc $?"; tail -3 ${S}_engine_tests.txt | cut -c1-300
  timeout 600 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"; cut -c1-200 ${S}_bench.json
fi
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ffn_ln_sk -s 4 -c 1 -f -o ${S}_ffn python tools/profile_ffn.py > ${S}_ncu_ffn.log 2>&1; echo "ncu ffn rc $?"

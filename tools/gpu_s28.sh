#!/bin/bash
# round 2, GPU session 28: stream-K FFN kernel -- parity, A/B timing against the split-tail and plain plans, smoke with the fine-tune step
mkdir -p gpurun_out
S=gpurun_out/r2s28
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k fused_ffn -x > ${S}_ffn_tests.txt 2>&1; rc=$?; echo "ffn tests rc $rc"; tail -15 ${S}_ffn_tests.txt | cut -c1-300
if [ $rc -eq 0 ]; then
  DTLR_TEST_HALF=f16 timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k fused_ffn -x > ${S}_ffn_tests_f16.txt 2>&1; echo "ffn tests f16 rc $?"; tail -3 ${S}_ffn_tests_f16.txt | cut -c1-300
  timeout 300 python tools/bench_ffn.py > ${S}_ffn_bench.txt 2>&1; cat ${S}_ffn_bench.txt
fi
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > ${S}_smoke.txt 2>&1; echo "smoke rc $?"; tail -8 ${S}_smoke.txt | cut -c1-300
if [ $rc -eq 0 ]; then
  timeout 900 python bench.py --steps 10 --warmup 3 > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"; cut -c1-600 ${S}_bench.json
fi

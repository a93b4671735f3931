#!/bin/bash
# round 2, GPU session 32: busy-polling barrier waits probe
mkdir -p gpurun_out
S=gpurun_out/r2s32
timeout 120 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k cta_pairs -x > ${S}_pair_tests.txt 2>&1; rc=$?; echo "pair tests rc $rc"; tail -30 ${S}_pair_tests.txt | cut -c1-400
if [ $rc -eq 0 ]; then
  DTLR_TEST_HALF=f16 timeout 120 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k cta_pairs -x > ${S}_pair_tests_f16.txt 2>&1; echo "pair tests f16 rc $?"; tail -3 ${S}_pair_tests_f16.txt | cut -c1-300
  FFN_PROBES=1 timeout 150 python tools/bench_ffn.py 58368 > ${S}_ffn_probes.txt 2>&1; cat ${S}_ffn_probes.txt
fi

#!/bin/bash
# round 2, GPU session 29: stream-K FFN on CTA pairs (cta_group::2) -- parity + A/B timing
mkdir -p gpurun_out
S=gpurun_out/r2s29
timeout 180 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k cta_pairs -x > ${S}_pair_tests.txt 2>&1; rc=$?; echo "pair tests rc $rc"; tail -30 ${S}_pair_tests.txt | cut -c1-400
if [ $rc -eq 0 ]; then
  DTLR_TEST_HALF=f16 timeout 180 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k cta_pairs -x > ${S}_pair_tests_f16.txt 2>&1; echo "pair tests f16 rc $?"; tail -3 ${S}_pair_tests_f16.txt | cut -c1-300
  timeout 300 python tools/bench_ffn.py > ${S}_ffn_bench.txt 2>&1; cat ${S}_ffn_bench.txt
  DTLR_DEBUG_FLAGS=1073741824 timeout 600 python bench.py --steps 10 --warmup 3 > ${S}_bench_pairs.json 2> ${S}_bench_pairs.err; echo "bench rc $?"; cut -c1-200 ${S}_bench_pairs.json
  timeout 600 python bench.py --steps 10 --warmup 3 > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"; cut -c1-200 ${S}_bench.json
fi

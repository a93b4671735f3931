#!/bin/bash
# round 2, GPU session 35: LayerNorm pass 2 interleaved with the next item's E1 steps
mkdir -p gpurun_out
S=gpurun_out/r2s35
timeout 120 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "cta_pairs or fused_ffn" -x > ${S}_tests.txt 2>&1; rc=$?; echo "tests rc $rc"; tail -5 ${S}_tests.txt | cut -c1-400
if [ $rc -eq 0 ]; then
  FFN_PROBES=1 timeout 150 python tools/bench_ffn.py 58368 > ${S}_ffn_probes.txt 2>&1; grep -v "SK " ${S}_ffn_probes.txt
  timeout 120 python tools/ffn_timeline.py > ${S}_ffn_timeline.txt 2>&1; echo rc $?; sed -n '1,25p' ${S}_ffn_timeline.txt | cut -c1-150; grep -A20 "epilogue warp 2" ${S}_ffn_timeline.txt | cut -c1-180; grep -A5 "final epilogues" ${S}_ffn_timeline.txt
fi

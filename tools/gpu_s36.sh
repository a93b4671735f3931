#!/bin/bash
# round 2, GPU session 36: stream-K FFN on CTA pairs as the default -- full GPU suite (both 16-bit aliases), bench line, ncu launch list of the step,
# ncu --set full of the new FFN kernel
mkdir -p gpurun_out
S=gpurun_out/r2s36
timeout 900 python -m pytest tests -q -m gpu -x > ${S}_suite.txt 2>&1; echo "suite rc $?"; tail -4 ${S}_suite.txt | cut -c1-300
DTLR_TEST_HALF=f16 timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_engine.py -q -m gpu -x > ${S}_suite_f16.txt 2>&1; echo "f16 gemm+engine rc $?"; tail -3 ${S}_suite_f16.txt | cut -c1-300
timeout 200 python tools/bench_ffn.py > ${S}_ffn_bench.txt 2>&1; cat ${S}_ffn_bench.txt
timeout 900 python bench.py --steps 10 --warmup 3 > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"; cut -c1-200 ${S}_bench.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${S}_launches.csv python tools/profile_step.py 2 > /dev/null 2>&1; echo "ncu list rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ffn_ln_sk -s 8 -c 1 -f -o ${S}_ffn python tools/profile_ffn.py > ${S}_ncu_ffn.log 2>&1; echo "ncu ffn rc $?"

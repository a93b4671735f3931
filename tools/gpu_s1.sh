#!/bin/bash
# round 2, GPU session 1: full parity suite (+ new bench-shape / HWDB / drop-in tests), grouped MSDA backward under the suite, bench line,
# fresh launch list + full ncu captures of the three hot kernels
mkdir -p gpurun_out
S=gpurun_out/r2s1
python -m pytest tests -m gpu -q -s -k "bench_shape or hwdb or reference_module" > ${S}_newtests.log 2>&1; echo "newtests rc $?"
python -m pytest tests -m gpu -x -q > ${S}_tests.log 2>&1; echo "tests rc $?"; tail -3 ${S}_tests.log
DTLR_DEBUG_FLAGS=65536 python -m pytest tests/test_gpu_msda.py tests/test_gpu_msda_vs_ref_cuda.py tests/test_gpu_dino_modules.py -m gpu -x -q > ${S}_tests_bwdgrouped.log 2>&1; echo "grouped rc $?"; tail -2 ${S}_tests_bwdgrouped.log
python bench.py --steps 10 --warmup 3 > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"; cat ${S}_bench.json | head -c 3000
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${S}_launches.csv python tools/profile_step.py 2 > ${S}_ll.log 2>&1; echo "launch list rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 4 -c 1 -f -o ${S}_msda python tools/profile_msda.py > ${S}_ncu_msda.log 2>&1; echo "ncu msda rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ffn_ln -s 4 -c 1 -f -o ${S}_ffn python tools/profile_ffn.py > ${S}_ncu_ffn.log 2>&1; echo "ncu ffn rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mha_flash -s 2 -c 1 -f -o ${S}_mha python tools/profile_attn.py hmma > ${S}_ncu_mha.log 2>&1; echo "ncu mha rc $?"
python tools/bench_msda_bwd.py > ${S}_bwd.log 2>&1; tail -5 ${S}_bwd.log
ls -la gpurun_out | tail -15

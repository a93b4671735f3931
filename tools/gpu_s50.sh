#!/bin/bash
# round 2, GPU session 50: state check -- smoke(), full bench line (all legs), reference arm, ncu launch list + ncu --set full of the attention kernel
mkdir -p gpurun_out
S=gpurun_out/r2s50
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > ${S}_smoke.txt 2>&1; echo "smoke rc $?"; tail -4 ${S}_smoke.txt | cut -c1-200
timeout 900 python bench.py > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"; cut -c1-200 ${S}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > ${S}_bench_ref.json 2> ${S}_bench_ref.err; echo "reference arm rc $?"; cut -c1-300 ${S}_bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${S}_launches.csv python tools/profile_step.py 2 > /dev/null 2>&1; echo "ncu list rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mha_tc2 -s 2 -c 1 -f -o ${S}_mha python tools/profile_attn.py tc > ${S}_ncu_mha.log 2>&1; echo "ncu mha rc $?"; tail -2 ${S}_ncu_mha.log

#!/bin/bash
# round 2, GPU session 62: ncu --set full of the split-precision tile kernel on the two FFN products (split3 / plain walk); fine-tune step timing re-check
mkdir -p gpurun_out
S=gpurun_out/r2s62
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 4 -c 2 -f -o ${S}_split_gemm python tools/profile_split_gemm.py > ${S}_ncu.log 2>&1; echo "ncu split3 rc $?"
timeout 200 ncu --set full --clock-control none -k regex:gemm_bf16_tcgen05 -s 4 -c 2 -f -o ${S}_split_gemm_plain python tools/profile_split_gemm.py plain > ${S}_ncu_plain.log 2>&1; echo "ncu plain rc $?"
timeout 300 python bench.py --no-cpu-baseline --no-gpu-reference --no-parity-mode > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s62_bench.json"))
print("value", d["value"], "train", d["train_step"].get("ms_per_step"))
PY

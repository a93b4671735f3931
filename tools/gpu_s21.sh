#!/bin/bash
# round 2, GPU session 21: backward-kernel tests, engine parity, true kernel durations of the native step (PDL off), bench line
mkdir -p gpurun_out
S=gpurun_out/r2s21
timeout 900 python -m pytest tests/test_gpu_train_kernels.py -q -m gpu > ${S}_kernels.txt 2>&1; echo "kernel tests rc $?"; tail -15 ${S}_kernels.txt
timeout 900 python -m pytest tests/test_gpu_train_engine.py -q -m gpu -s > ${S}_engine.txt 2>&1; echo "engine tests rc $?"; grep "worst\|bf16 loss\|passed\|failed" ${S}_engine.txt | cut -c1-300
DTLR_DEBUG_FLAGS=64 DTLR_TRAIN_PROFILE=1 timeout 900 python tools/bench_train_native.py 32 bf16 > ${S}_train_nopdl.txt 2>&1; echo "profile rc $?"; head -64 ${S}_train_nopdl.txt | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 --train-ab > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s21_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d.get("e2e"), d.get("train_step"))
PY

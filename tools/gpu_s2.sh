#!/bin/bash
# round 2, GPU session 2: new kernels (fused CTC loss, select / PostProcess / NMS) + full suite + bench
mkdir -p gpurun_out
S=gpurun_out/r2s2
python -m pytest tests/test_gpu_ctc_loss.py tests/test_gpu_select.py -m gpu -q -s > ${S}_new.log 2>&1; echo "new rc $?"; tail -25 ${S}_new.log
python -m pytest tests -m gpu -q > ${S}_tests.log 2>&1; echo "tests rc $?"; tail -8 ${S}_tests.log
python bench.py --steps 10 --warmup 3 > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s2_bench.json'))
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['train_step'], d['gpu_reference'])
PY

#!/bin/bash
# round 2, GPU session 25: full GPU suite with the native fine-tune step, MSDA-backward group-size A/B inside the step, bench line
mkdir -p gpurun_out
S=gpurun_out/r2s25
timeout 1500 python -m pytest tests -q -m gpu -x > ${S}_suite.txt 2>&1; echo "suite rc $?"; tail -5 ${S}_suite.txt | cut -c1-300
for v in 0 131072; do
  DTLR_DEBUG_FLAGS=$v timeout 600 python tools/bench_train_native.py 32 bf16 > ${S}_train_$v.txt 2>&1; echo "flags $v: $(grep variant ${S}_train_$v.txt)"
done
timeout 900 python bench.py --steps 10 --warmup 3 --train-ab > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s25_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], {k: v for k, v in d["train_step"].items() if k not in ("what",)})
PY

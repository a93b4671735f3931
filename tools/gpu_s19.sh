#!/bin/bash
# round 2, GPU session 19: where the fine-tune step goes (torch.profiler), LayerNorm two-rows-per-warp A/B
mkdir -p gpurun_out
S=gpurun_out/r2s19
timeout 600 python tools/profile_train_step.py 32 > ${S}_train_profile.txt 2> ${S}_train_profile.err; echo "profile rc $?"
head -50 ${S}_train_profile.txt
for v in 0 67108864; do
  DTLR_DEBUG_FLAGS=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench_$v.json 2> ${S}_bench_$v.err; echo "bench flags $v rc $?"
done
python - <<'PY'
import json
for t in ("0", "67108864"):
    try:
        d = json.load(open("gpurun_out/r2s19_bench_%s.json" % t))
        print(t, {k: d[k] for k in ("value", "ms_per_step")})
    except Exception as e:
        print(t, "failed", e)
PY

#!/bin/bash
# round 2, GPU session 54 (2 GPUs): bench under torchrun N=2 with the tcgen05 attention default + CTA-pair FFN
mkdir -p gpurun_out
S=gpurun_out/r2s54
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > ${S}_bench_n2.json 2> ${S}_bench_n2.err; echo "bench n2 rc $?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2s54_bench_n2.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"], json.dumps({k: v for k, v in d.get("train_step").items() if k != "what"})[:600])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2s54_bench_n2.err").read()[-2000:])
PY

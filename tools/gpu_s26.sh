#!/bin/bash
# round 2, GPU session 26 (2 GPUs): bench under torchrun N=2 with the native fine-tune step (segmented async all-reduce) + torch DDP A/B
mkdir -p gpurun_out
S=gpurun_out/r2s26
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference --train-ab > ${S}_bench_n2.json 2> ${S}_bench_n2.err; echo "bench n2 rc $?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2s26_bench_n2.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, json.dumps({k: v for k, v in d.get("train_step").items() if k != "what"})[:900])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2s26_bench_n2.err").read()[-2000:])
PY

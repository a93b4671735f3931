#!/bin/bash
# round 2, GPU session 23 (2 GPUs): native fine-tune step under torchrun N=2 (segmented async all-reduce), single-GPU tests + timing
mkdir -p gpurun_out
S=gpurun_out/r2s23
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_engine.py -q -m gpu -s > ${S}_tests.txt 2>&1; echo "tests rc $?"; grep "worst\|bf16 loss\|passed\|failed\|Error" ${S}_tests.txt | cut -c1-300
timeout 600 python tools/bench_train_native.py 32 bf16 > ${S}_train.txt 2>&1; echo "timing rc $?"; grep variant ${S}_train.txt
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference --train-ab > ${S}_bench_n2.json 2> ${S}_bench_n2.err; echo "bench n2 rc $?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2s23_bench_n2.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, json.dumps(d.get("train_step"))[:900])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2s23_bench_n2.err").read()[-2000:])
PY

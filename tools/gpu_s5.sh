#!/bin/bash
# round 2, GPU session 5 (2 GPUs): the bench line under torchrun at N = 2 (inference shards + the DDP fine-tune step), reference arm at N = 2
mkdir -p gpurun_out
S=gpurun_out/r2s5
DTLR_TEST_HALF=f16 timeout 300 python -m pytest tests/test_gpu_attention.py tests/test_gpu_msda.py -m gpu -q > ${S}_f16_attn.log 2>&1; echo "f16 attention/msda rc $?"; tail -3 ${S}_f16_attn.log
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > ${S}_bench_n2.json 2> ${S}_bench_n2.err; echo "bench N=2 rc $?"
wc -l ${S}_bench_n2.json; grep -c "NCCL INFO" ${S}_bench_n2.err; grep -m3 "nranks\|NVLS\|Connected all" ${S}_bench_n2.err | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > ${S}_ref_n2.json 2> ${S}_ref_n2.err; echo "reference N=2 rc $?"; cat ${S}_ref_n2.json | cut -c1-400
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s5_bench_n2.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "dtype", "gpu_launches")}, "e2e", d["e2e"]["value"], "train", d.get("train_step"))
PY

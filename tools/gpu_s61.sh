#!/bin/bash
# round 2, GPU session 61: final-state check -- full GPU suite in both 16-bit flavours, smoke(), default bench line (all legs incl. parity_mode), reference arm
mkdir -p gpurun_out
S=gpurun_out/r2s61
timeout 400 python -m pytest tests -m gpu -x -q > ${S}_suite.txt 2>&1; echo "suite rc $?"; tail -2 ${S}_suite.txt
DTLR_TEST_HALF=f16 timeout 400 python -m pytest tests -m gpu -x -q > ${S}_suite_f16.txt 2>&1; echo "suite f16 rc $?"; tail -2 ${S}_suite_f16.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${S}_smoke.txt 2>&1; echo "smoke rc $?"; grep -a "smoke" ${S}_smoke.txt | cut -c1-200
timeout 600 python bench.py > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2s61_bench.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity_mode", json.dumps(d.get("parity_mode"))[:700])
    print("cpu", d["cpu_baseline"]["value"], "gpu_ref", d["gpu_reference"].get("value"), "train", d["train_step"].get("ms_per_step"), "ffn frac", d["roofline"]["frac"], "msda frac", d["roofline_msda"]["frac"], "launches", d["gpu_launches"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2s61_bench.err").read()[-1500:])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > ${S}_bench_ref.json 2> ${S}_bench_ref.err; echo "reference arm rc $?"; cut -c1-200 ${S}_bench_ref.json

#!/bin/bash
# round 2, GPU session 56: final-state check (full GPU suite, default bench line with the new parity_mode leg), then compute-sanitizer
# memcheck over smoke() (every kernel of the forward in three dtypes + the native fine-tune step) and the attention / GEMM+FFN edge-shape tests
mkdir -p gpurun_out
S=gpurun_out/r2s56
timeout 300 python -m pytest tests -m gpu -x -q > ${S}_suite.txt 2>&1; echo "suite rc $?"; tail -2 ${S}_suite.txt
timeout 400 python bench.py > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"; cut -c1-160 ${S}_bench.json
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2s56_bench.json"))
    print("e2e", d["e2e"]["value"], "parity_mode", d.get("parity_mode"), "cpu", d["cpu_baseline"]["value"], "gpu_ref", d["gpu_reference"].get("value"),
          "train", d["train_step"].get("ms_per_step"), "ffn frac", d["roofline"]["frac"], "msda frac", d["roofline_msda"]["frac"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2s56_bench.err").read()[-1500:])
PY
SAN="compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 7 --print-limit 20"
timeout 300 $SAN python -c "import __graft_entry__ as g; g.smoke()" > ${S}_memcheck_smoke.txt 2>&1; echo "memcheck smoke rc $?"; tail -3 ${S}_memcheck_smoke.txt | cut -c1-200
timeout 110 $SAN python -m pytest tests/test_gpu_attention.py -x -q > ${S}_memcheck_attn.txt 2>&1; echo "memcheck attention rc $?"; tail -3 ${S}_memcheck_attn.txt | cut -c1-200
timeout 110 $SAN python -m pytest tests/test_gpu_gemm.py -x -q > ${S}_memcheck_gemm.txt 2>&1; echo "memcheck gemm rc $?"; tail -3 ${S}_memcheck_gemm.txt | cut -c1-200

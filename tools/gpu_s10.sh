#!/bin/bash
# round 2, GPU session 10: TMA staging of the MSDA value slab (A/B vs cp.async), suite in both flavours, full bench line, fresh ncu of msda / ffn / mha
mkdir -p gpurun_out
S=gpurun_out/r2s10
timeout 300 python -m pytest tests/test_gpu_msda.py tests/test_gpu_msda_vs_ref_cuda.py tests/test_reference_ops_dropin.py -m gpu -q > ${S}_msda_tests.log 2>&1; echo "msda tests rc $?"; tail -3 ${S}_msda_tests.log
timeout 600 python -m pytest tests -m gpu -q > ${S}_tests.log 2>&1; echo "tests rc $?"; tail -3 ${S}_tests.log
DTLR_TEST_HALF=f16 timeout 600 python -m pytest tests -m gpu -q > ${S}_tests_f16.log 2>&1; echo "f16-aliased tests rc $?"; tail -3 ${S}_tests_f16.log
timeout 300 python tools/bench_msda.py > ${S}_msda_tma.log 2>&1; tail -6 ${S}_msda_tma.log
DTLR_DEBUG_FLAGS=2097152 timeout 300 python tools/bench_msda.py > ${S}_msda_cpasync.log 2>&1; tail -6 ${S}_msda_cpasync.log
timeout 900 python bench.py --steps 10 --warmup 3 > ${S}_bench.json 2> ${S}_bench.err; echo "bench rc $?"
DTLR_DEBUG_FLAGS=2097152 timeout 900 python bench.py --steps 10 --warmup 3 --no-train-step --no-gpu-reference --no-cpu-baseline > ${S}_bench_cpasync.json 2> ${S}_bench_cpasync.err
python - <<'PY'
import json
for t in ("bench", "bench_cpasync"):
    try:
        d = json.load(open("gpurun_out/r2s10_%s.json" % t))
        print(t, {k: d[k] for k in ("value", "ms_per_step", "dtype", "gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e_u8"]["value"], "ffn", d["roofline"]["us_per_launch"], d["roofline"]["frac"], "msda", d["roofline_msda"]["us_per_launch"], d["roofline_msda"]["frac"], "train", (d.get("train_step") or {}).get("ms_per_step"), "gpu_ref", (d.get("gpu_reference") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(t, "failed", e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 4 -c 1 -f -o ${S}_msda python tools/profile_msda.py > ${S}_ncu_msda.log 2>&1; echo "ncu msda rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ffn_ln_tcgen05 -s 8 -c 1 -f -o ${S}_ffn python tools/profile_ffn.py > ${S}_ncu_ffn.log 2>&1; echo "ncu ffn rc $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mha_flash -s 2 -c 1 -f -o ${S}_mha python tools/profile_attn.py flash > ${S}_ncu_mha.log 2>&1; echo "ncu mha rc $?"

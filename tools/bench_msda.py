"""Micro-benchmark of the MSDA core at BASELINE config-2 size (B=64, S=912, Lq=912|900, M=8, D=32, L=4, P=4).
CUDA events, warm-up, working set (> 126 MB L2) larger than L2.  Prints one JSON line per variant:
un-fused core (fp32 / bf16 values) and the fused-prologue call the engine makes (bf16 values + bf16 projection rows),
each bf16 variant on the tensor-core gather kernel and, with dtlr_debug_flags(16), on the SIMT kernel."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import _lib, msda  # noqa: E402

PEAK = 6580.6
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

S, M, D, L, P = 912, 8, 32, 4, 4


def _time(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / iters


def _levels():
    shapes = torch.tensor([(5, 128), (3, 64), (2, 32), (1, 16)], dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    return msda._host_levels(shapes.cuda(), lsi.cuda())


def run(dtype, B=64, Lq=912, iters=20, kernel="default"):
    sh, ls, n = _levels()
    g = torch.Generator(device="cuda").manual_seed(0)
    value = torch.randn(B, S, M, D, device="cuda", generator=g).to(dtype)
    loc = torch.rand(B, Lq, M, L, P, 2, device="cuda", generator=g)
    w = torch.softmax(torch.randn(B, Lq, M, L * P, device="cuda", generator=g), -1).view(B, Lq, M, L, P)
    out = torch.empty(B, Lq, M * D, device="cuda", dtype=dtype)
    _lib.lib().dtlr_debug_flags(16 if kernel == "simt" else 0)
    us = _time(lambda: msda.msda_forward_raw(value, sh, ls, n, loc, w, out), iters)
    _lib.lib().dtlr_debug_flags(0)
    es = value.element_size()
    alg = B * (S * M * D * es + Lq * M * L * P * 2 * 4 + Lq * M * L * P * 4 + Lq * M * D * es) + 96
    print(json.dumps({"kernel": "msda_fwd", "variant": kernel, "dtype": str(dtype), "B": B, "Lq": Lq, "us": round(us, 2),
                      "alg_MB": round(alg / 1e6, 2), "GBs": round(alg / us / 1e3, 1),
                      "frac_hbm": round(alg / us / 1e3 / PEAK, 3)}), flush=True)


def run_fused(B=64, Lq=912, ref_dim=2, iters=20, kernel="default"):
    """the engine's call: bf16 values (column block of the value projection), bf16 offsets|logits rows, fp32 reference points"""
    sh, ls, n = _levels()
    g = torch.Generator(device="cuda").manual_seed(0)
    value = torch.randn(B, S, M, D, device="cuda", generator=g).bfloat16()
    proj = torch.randn(B * Lq, 384, device="cuda", generator=g).bfloat16()
    ref = torch.rand(B * Lq, ref_dim, device="cuda", generator=g)
    if ref_dim == 4:
        ref[:, 2:] *= 0.3
    vr = torch.ones(B, L, 2, device="cuda")
    out = torch.empty(B, Lq, M * D, device="cuda", dtype=torch.bfloat16)
    _lib.lib().dtlr_debug_flags({"simt": 16, "qu1": 4194304, "qu2": 8388608, "cpasync": 2097152, "cpasync_qu1": 2097152 + 4194304}.get(kernel, 0))
    us = _time(lambda: msda.msda_forward_fused(value, sh, ls, n, proj, ref, vr, Lq, P, out), iters)
    _lib.lib().dtlr_debug_flags(0)
    # algorithmic bytes (SURVEY 8d with bf16 value/out, and the projection rows + reference points instead of loc / weights)
    alg = B * (S * M * D * 2 + Lq * 384 * 2 + Lq * ref_dim * 4 + Lq * M * D * 2) + 96
    print(json.dumps({"kernel": "msda_fwd_fused", "variant": kernel, "ref_dim": ref_dim, "B": B, "Lq": Lq, "us": round(us, 2),
                      "alg_MB": round(alg / 1e6, 2), "GBs": round(alg / us / 1e3, 1),
                      "frac_hbm": round(alg / us / 1e3 / PEAK, 3)}), flush=True)


if __name__ == "__main__":
    for kern in ("default", "qu1", "qu2", "cpasync", "cpasync_qu1"):      # phase-2 unroll 4 (default) / 1 / 2; slab by TMA (default) / cp.async
        run_fused(Lq=912, ref_dim=2, kernel=kern)
        run_fused(Lq=900, ref_dim=4, kernel=kern)
    for kern in ("default", "simt"):
        run_fused(Lq=912, ref_dim=2, kernel=kern)
        run_fused(Lq=900, ref_dim=4, kernel=kern)
        run(torch.bfloat16, Lq=912, kernel=kern)
    run(torch.float32, Lq=912)
    run_fused(B=8, Lq=912, ref_dim=2)
    run(torch.bfloat16, B=8)

"""Fused FFN block (dtlr_ffn_ln_ws) at the bench shapes, timed as a CUDA graph of N back-to-back calls on rotating inputs
(> L2 working set), per plan: stream-K on CTA pairs (default) / on single CTAs (dtlr_debug_flags 1073741824) / full rounds + split
tail (536870912) / plain (262144).
python tools/bench_ffn.py [M ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import _lib, ops  # noqa: E402

hid = 2048
dt = torch.float16 if os.environ.get("DTLR_TEST_HALF", "f16") == "f16" else torch.bfloat16
Ms = [int(a) for a in sys.argv[1:]] or [58368, 57600]
peak = 1383.8
for M in Ms:
    g = torch.Generator(device="cuda").manual_seed(0)
    x = [torch.randn(M, 256, device="cuda", generator=g).to(dt) for _ in range(6)]
    w1 = (torch.randn(hid, 256, device="cuda", generator=g) / 16).to(dt)
    w2 = (torch.randn(256, hid, device="cuda", generator=g) / 45).to(dt)
    b1 = torch.randn(hid, device="cuda", generator=g) * 0.1
    b2 = torch.randn(256, device="cuda", generator=g) * 0.1
    gm = torch.ones(256, device="cuda")
    bt = torch.zeros(256, device="cuda")
    _lib.set_flavor(dt)
    variants = [("stream-K, CTA pairs (default)", 0), ("stream-K, single CTAs", 1073741824), ("full rounds + split tail", 536870912), ("plain", 262144)]
    for name, flags in variants:
        _lib.lib().dtlr_debug_flags(flags)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for i in range(3):
                ops.ffn_ln(x[i % 6], w1, b1, w2, b2, gm, bt)
        torch.cuda.synchronize()
        N = 24
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(N):
                ops.ffn_ln(x[i % 6], w1, b1, w2, b2, gm, bt)
        graph.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / N)
        tf = 4.0 * M * hid * 256 / (best * 1e-3) / 1e12
        print("M %d %-44s plan %d: %.1f us per block, %.0f TFLOP/s = %.3f of the sustained peak (%.1f)" % (
            M, name, _lib.lib().dtlr_ffn_plan(M, hid), best * 1e3, tf, tf / peak, peak), flush=True)
        _lib.lib().dtlr_debug_flags(0)

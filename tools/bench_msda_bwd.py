"""Micro-benchmark of the MSDA backward (reference row a10) at the BASELINE config-5 per-GPU size (B=32, S=912, Lq=900, M=8,
D=32, L=4, P=4, fp32): the D=32 fast kernel vs the generic one-warp-per-item kernel (dtlr_debug_flags(8192)).
CUDA events around 10 calls after warm-up (each call = memset of grad_value + kernel).  Algorithmic bytes: forward
bytes + grad_out read + grad_value / grad_loc / grad_attn written."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import _lib, msda  # noqa: E402

S, M, D, L, P = 912, 8, 32, 4, 4


def run(B, Lq, flags):
    shapes = torch.tensor([(5, 128), (3, 64), (2, 32), (1, 16)], dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    g = torch.Generator(device="cuda").manual_seed(0)
    value = torch.randn(B, S, M, D, device="cuda", generator=g)
    loc = torch.rand(B, Lq, M, L, P, 2, device="cuda", generator=g)
    w = torch.softmax(torch.randn(B, Lq, M, L * P, device="cuda", generator=g), -1).view(B, Lq, M, L, P)
    go = torch.randn(B, Lq, M * D, device="cuda", generator=g)
    sh, ls = shapes.cuda(), lsi.cuda()
    _lib.lib().dtlr_debug_flags(flags)
    fn = lambda: msda.ms_deform_attn_backward(value, sh, ls, loc, w, go, 64)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    _lib.lib().dtlr_debug_flags(0)
    us = e0.elapsed_time(e1) * 100
    fwd_bytes = B * (S * M * D * 4 + Lq * M * L * P * 12 + Lq * M * D * 4)
    bwd_bytes = fwd_bytes + B * (S * M * D * 4 + Lq * M * L * P * 12)
    print(json.dumps({"op": "msda_backward", "kernel": {65536: "d32_one_load_in_flight", 8192: "generic", 0: "d32_grouped_loads_4 (default)", 131072: "d32_grouped_loads_8"}[flags], "B": B, "Lq": Lq, "us": round(us, 1),
                      "algorithmic_GBps": round(bwd_bytes / us / 1e3, 1)}), flush=True)


if __name__ == "__main__":
    import sys as _s
    for B, Lq in ((32, 900),) if "quick" in _s.argv else ((32, 900), (32, 1082), (64, 912)):
        run(B, Lq, 0)
        run(B, Lq, 65536)       # the round-1 default: one value load in flight per warp
        run(B, Lq, 131072)      # 8 in flight, 128-thread CTAs (measured slower: 1054 vs 855 us at B = 64)
        run(B, Lq, 8192)

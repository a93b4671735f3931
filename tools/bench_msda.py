"""Micro-benchmark of the MSDA core at BASELINE config-2 size (B=64, S=912, Lq=912|900, M=8, D=32, L=4, P=4).
CUDA events, warm-up, working set (> 126 MB L2) larger than L2.  Prints one JSON line per variant."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import msda  # noqa: E402

PEAK = 6580.6
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def run(dtype, B=64, Lq=912, iters=20):
    shapes = torch.tensor([(5, 128), (3, 64), (2, 32), (1, 16)], dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S, M, D, L, P = 912, 8, 32, 4, 4
    g = torch.Generator(device="cuda").manual_seed(0)
    value = torch.randn(B, S, M, D, device="cuda", generator=g).to(dtype)
    loc = torch.rand(B, Lq, M, L, P, 2, device="cuda", generator=g)
    w = torch.softmax(torch.randn(B, Lq, M, L * P, device="cuda", generator=g), -1).view(B, Lq, M, L, P)
    sh, ls, n = msda._host_levels(shapes.cuda(), lsi.cuda())
    out = torch.empty(B, Lq, M * D, device="cuda", dtype=dtype)
    for _ in range(5):
        msda.msda_forward_raw(value, sh, ls, n, loc, w, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        msda.msda_forward_raw(value, sh, ls, n, loc, w, out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / iters
    es = value.element_size()
    alg = B * (S * M * D * es + Lq * M * L * P * 2 * 4 + Lq * M * L * P * 4 + Lq * M * D * es) + 96
    print(json.dumps({"kernel": "msda_fwd", "dtype": str(dtype), "B": B, "Lq": Lq, "us": round(us, 2),
                      "alg_MB": round(alg / 1e6, 2), "GBs": round(alg / us / 1e3, 1),
                      "frac_hbm": round(alg / us / 1e3 / PEAK, 3)}))


if __name__ == "__main__":
    for dt in (torch.float32, torch.bfloat16):
        for Lq in (912, 900):
            run(dt, Lq=Lq)
    run(torch.float32, B=8)
    run(torch.bfloat16, B=8)

"""The deformable-attention core against the reference's own CUDA kernels recompiled for sm_100a (oracle/_ref, built by
oracle/build_ref_cuda.py) on the same B200, BASELINE config-2 / config-5 sizes.  CUDA events, 20 calls after 5 warm-ups, inputs of
one call (value + loc + w + out = 209 MB at B = 64 fp32) larger than the 126 MB L2.  One JSON line per kernel."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # repo root
from dtlr_b200 import msda  # noqa: E402
from oracle import build_ref_cuda  # noqa: E402   (dev tooling: the reference kernel is the measured baseline here)

S, M, D, L, P = 912, 8, 32, 4, 4


def timed(fn, iters=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    ref = build_ref_cuda.load()
    if ref is None:
        print(json.dumps({"error": "oracle/_ref/MultiScaleDeformableAttention.so not built"}))
        return
    shp = torch.tensor([(5, 128), (3, 64), (2, 32), (1, 16)], dtype=torch.long, device="cuda")
    lsi = torch.cat((shp.new_zeros((1,)), shp.prod(1).cumsum(0)[:-1]))
    for B, Lq in ((64, 912), (32, 900)):
        g = torch.Generator(device="cuda").manual_seed(0)
        value = torch.randn(B, S, M, D, device="cuda", generator=g)
        loc = torch.rand(B, Lq, M, L, P, 2, device="cuda", generator=g)
        w = torch.softmax(torch.randn(B, Lq, M, L * P, device="cuda", generator=g), -1).view(B, Lq, M, L, P)
        go = torch.randn(B, Lq, M * D, device="cuda", generator=g)
        v16 = value.bfloat16()
        rows = {
            "forward fp32: reference CUDA kernel (ms_deformable_im2col_gpu_kernel)": lambda: ref.ms_deform_attn_forward(value, shp, lsi, loc, w, B),
            "forward fp32: dtlr_msda_forward (slab kernel)": lambda: msda.ms_deform_attn_forward(value, shp, lsi, loc, w, 64),
            "forward bf16 values: dtlr_msda_forward (tensor-core gather)": lambda: msda.ms_deform_attn_forward(v16, shp, lsi, loc, w, 64),
            "backward fp32: reference CUDA kernel (col2im, blocksize-aware reduce)": lambda: ref.ms_deform_attn_backward(value, shp, lsi, loc, w, go, B),
            "backward fp32: dtlr_msda_backward (d32 fast kernel)": lambda: msda.ms_deform_attn_backward(value, shp, lsi, loc, w, go, 64),
        }
        for name, fn in rows.items():
            print(json.dumps({"B": B, "Lq": Lq, "kernel": name, "us": round(timed(fn), 1)}), flush=True)


if __name__ == "__main__":
    main()

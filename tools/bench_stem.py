"""Stem convolution (7x7 / 2, 3 -> 64, FrozenBN + ReLU folded) at B=64 x 3x40x1024: direct FFMA kernel vs im2col + tcgen05 GEMM.
CUDA-graph timing (no host launch gaps)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import ops
from gemm_probe_util import timeit

B, H, W = 64, 40, 1024
x = torch.randn(B, 3, H, W, device="cuda")
w = torch.randn(64, 3, 7, 7, device="cuda") * 0.05
bias = torch.randn(64, device="cuda")
w_direct = w.permute(2, 3, 1, 0).contiguous()
wk = torch.nn.functional.pad(w.permute(0, 2, 3, 1).reshape(64, 147), (0, 5)).bfloat16().contiguous()     # [64, 152], K order (kh, kw, cin)

d = timeit(lambda i: ops.stem_conv(x, w_direct, bias, B, H, W, torch.bfloat16))


def two(i):
    col, Ho, Wo = ops.im2col(x, B, H, W, 3, 7, 7, 2, 3, torch.bfloat16, nchw_input=True, ldo=152)
    return ops.gemm(col, wk, bias, relu=1)


t = timeit(two)
c = timeit(lambda i: ops.im2col(x, B, H, W, 3, 7, 7, 2, 3, torch.bfloat16, nchw_input=True, ldo=152))
y1 = ops.stem_conv(x, w_direct, bias, B, H, W, torch.bfloat16)[0].float()
y2 = two(0).float()
print("stem: direct %.1f us, im2col+gemm %.1f us (im2col alone %.1f), max diff %.3e (ref max %.2f)" % (d, t, c, (y1 - y2).abs().max().item(), y1.abs().max().item()))

"""The split-precision products of one FFN block at the bench shape (M = 58368) for ncu: linear1 (fp32 A -> split pass -> N = 2048 product with
the split-output epilogue) and linear2 (split A, K = 2048 -> fp32 + residual), hi / lo tiles loaded once (split3) -- and the plain 3K walk
with argument `plain`."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import engine, ops
if len(sys.argv) > 1 and sys.argv[1] == "plain":
    ops.SPLIT3_LOADS = False
M, half = 58368, torch.float16
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, 256, device="cuda", generator=g)
w1 = engine._split_w(torch.randn(2048, 256, device="cuda", generator=g) / 16, half)
w2 = engine._split_w(torch.randn(256, 2048, device="cuda", generator=g) / 45, half)
b1, b2 = torch.randn(2048, device="cuda", generator=g), torch.randn(256, device="cuda", generator=g)
for _ in range(4):
    h = ops.gemm(x, w1, b1, relu=1, out_dtype=ops.SPLIT)
    y = ops.gemm(h, w2, b2, residual=x, out_dtype=torch.float32, split3=True)
torch.cuda.synchronize()

"""The K = 256 projection kernels of a transformer layer at the bench shape (M = 58368 rows) for ncu:
weight-stationary GEMM (N = 256 and N = 384, with / without residual), add_layernorm256, the fused box head."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import ops
M, dt = 58368, torch.float16
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, 256, device="cuda", generator=g).to(dt)
res = torch.randn(M, 256, device="cuda", generator=g).to(dt)
w = (torch.randn(256, 256, device="cuda", generator=g) / 16).to(dt)
w384 = (torch.randn(384, 256, device="cuda", generator=g) / 16).to(dt)
b = torch.randn(256, device="cuda", generator=g)
b384 = torch.randn(384, device="cuda", generator=g)
gm, bt = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
w3 = torch.randn(4, 256, device="cuda", generator=g) * 0.05
b3 = torch.zeros(4, device="cuda")
ref = torch.rand(M, 4, device="cuda", generator=g)
for _ in range(4):
    y = ops.gemm(x, w, b)
    y2 = ops.gemm(x, w, b, residual=res)
    y3 = ops.gemm(x, w384, b384)
    z = ops.add_layernorm(y2, None, gm, bt)
    z2, z3 = ops.add_layernorm(y2, None, gm, bt, add2=res)
    h = ops.mlp_head(x, (w, b), (w, b), w3, b3, ref)
    a = ops.add(x, res)
torch.cuda.synchronize()

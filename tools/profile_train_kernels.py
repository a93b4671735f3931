"""ncu target: the new backward kernels at the fine-tune step's shapes (32 lines): wgrad (256x256, 2048x256 FFN), flash attention forward /
dQ / dKV passes, LayerNorm backward, column sums.  Usage: ncu ... python tools/profile_train_kernels.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import train_ops as K

dev = "cuda"
M = 32 * 912
Mq = 32 * 1014
torch.manual_seed(0)
x = torch.randn(M, 256, device=dev).bfloat16()
dy = torch.randn(M, 256, device=dev).bfloat16()
h = torch.randn(M, 2048, device=dev).bfloat16()
g1 = torch.zeros(256, 256, device=dev)
g2 = torch.zeros(2048, 256, device=dev)
g3 = torch.zeros(256, 2048, device=dev)
qk = torch.randn(Mq, 512, device=dev).bfloat16()
v = torch.randn(Mq, 256, device=dev).bfloat16()
do = torch.randn(Mq, 256, device=dev).bfloat16()
mask = torch.zeros(1014, 1014, dtype=torch.bool, device=dev)
mask[114:, :114] = True
mobj = K.make_mask(mask)
z = torch.randn(M, 256, device=dev).bfloat16()
d32 = torch.randn(M, 256, device=dev)
gam = torch.ones(256, device=dev)
dg, db = torch.zeros(256, device=dev), torch.zeros(256, device=dev)
bias = torch.zeros(2048, device=dev)
for it in range(3):
    K.wgrad(dy, x, g1)
    K.wgrad(h, x, g2)
    K.wgrad(dy, h, g3)
    att, ctx = K.sa_forward(qk, v, mobj, 32, 1014, 8)
    K.sa_backward(ctx, do)
    K.layernorm_bwd(z, d32, None, gam, dg, db)
    K.colsum(h, bias)
torch.cuda.synchronize()
print("ok")

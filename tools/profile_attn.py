"""A few decoder self-attention calls on the tcgen05 path for ncu:
ncu --set full --clock-control none --import-source on -k regex:mha_tc2 -s 2 -c 1 -o gpurun_out/mha_tc2 python tools/profile_attn.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import ops
ops.ATTN_IMPL = sys.argv[1] if len(sys.argv) > 1 else "tc"
B, Q, heads, d = 64, 900, 8, 256
qk = torch.randn(B * Q, 2 * d, device="cuda").bfloat16()
v = torch.randn(B * Q, d, device="cuda").bfloat16()
for _ in range(4):
    ops.mha_self_attention(qk, d, v, None, B, Q, heads, 32)
torch.cuda.synchronize()

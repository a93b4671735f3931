"""One fused FFN block call at BASELINE config-2 size (M = 64 x 912 rows, hidden 2048) for ncu (4 warm-up calls first).
ncu --set full --clock-control none --import-source on -k regex:ffn_ln -s 4 -c 1 -o gpurun_out/ffn python tools/profile_ffn.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import ops  # noqa: E402

M, hid = 58368, 2048
g = torch.Generator(device="cuda").manual_seed(0)
x = [torch.randn(M, 256, device="cuda", generator=g).bfloat16() for _ in range(3)]
w1 = (torch.randn(hid, 256, device="cuda", generator=g) / 16).bfloat16()
w2 = (torch.randn(256, hid, device="cuda", generator=g) / 45).bfloat16()
b1 = torch.randn(hid, device="cuda", generator=g) * 0.1
b2 = torch.randn(256, device="cuda", generator=g) * 0.1
gm = torch.ones(256, device="cuda")
bt = torch.zeros(256, device="cuda")
for i in range(6):
    ops.ffn_ln(x[i % 3], w1, b1, w2, b2, gm, bt)
torch.cuda.synchronize()

"""One MSDA backward call at the BASELINE config-5 per-GPU size (B=32, Lq=900, fp32) for ncu (3 warm-up calls first).
ncu --set full --clock-control none --import-source on -k regex:msda_bwd -s 3 -c 1 -o gpurun_out/msda_bwd python tools/profile_msda_bwd.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import _lib, msda  # noqa: E402

B, Lq, S, M, D, L, P = 32, 900, 912, 8, 32, 4, 4
shapes = torch.tensor([(5, 128), (3, 64), (2, 32), (1, 16)], dtype=torch.long)
lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
g = torch.Generator(device="cuda").manual_seed(0)
value = torch.randn(B, S, M, D, device="cuda", generator=g)
loc = torch.rand(B, Lq, M, L, P, 2, device="cuda", generator=g)
w = torch.softmax(torch.randn(B, Lq, M, L * P, device="cuda", generator=g), -1).view(B, Lq, M, L, P)
go = torch.randn(B, Lq, M * D, device="cuda", generator=g)
if len(sys.argv) > 1 and sys.argv[1] == "generic":
    _lib.lib().dtlr_debug_flags(8192)
for _ in range(4):
    msda.ms_deform_attn_backward(value, shapes.cuda(), lsi.cuda(), loc, w, go, 64)
torch.cuda.synchronize()

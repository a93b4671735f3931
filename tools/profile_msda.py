"""One fused MSDA call at BASELINE config-2 size for ncu (4 warm-up calls, then the captured ones).
ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 4 -c 1 -o gpurun_out/msda python tools/profile_msda.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import _lib, msda  # noqa: E402

B, Lq, S, M, D, L, P = 64, 912, 912, 8, 32, 4, 4
shapes = torch.tensor([(5, 128), (3, 64), (2, 32), (1, 16)], dtype=torch.long)
lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
sh, ls, n = msda._host_levels(shapes.cuda(), lsi.cuda())
g = torch.Generator(device="cuda").manual_seed(0)
value = torch.randn(B, S, M, D, device="cuda", generator=g).bfloat16()
proj = torch.randn(B * Lq, 384, device="cuda", generator=g).bfloat16()
ref = torch.rand(B * Lq, 2, device="cuda", generator=g)
vr = torch.ones(B, L, 2, device="cuda")
out = torch.empty(B, Lq, M * D, device="cuda", dtype=torch.bfloat16)
if len(sys.argv) > 1 and sys.argv[1] == "simt":
    _lib.lib().dtlr_debug_flags(16)
for _ in range(6):
    msda.msda_forward_fused(value, sh, ls, n, proj, ref, vr, Lq, P, out)
torch.cuda.synchronize()

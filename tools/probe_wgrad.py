"""Which (LBO, SBO) assignment of the MN-major SWIZZLE_128B shared-memory descriptor is right for dtlr_wgrad's tcgen05 kernel:
runs the default and the swapped variant (dtlr_debug_flags 134217728) against fp64 and prints both errors."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import _lib as L, train_ops as K

for flag in (0, 134217728):
    for rows, N, Kd in ((4096, 256, 256), (333, 166, 256), (2000, 128, 512)):
        dy = torch.randn(rows, (N + 7) // 8 * 8, device="cuda").bfloat16()
        x = torch.randn(rows, Kd, device="cuda").bfloat16()
        gw = torch.zeros(N, Kd, device="cuda")
        L.set_flavor("bf16")
        L.lib().dtlr_debug_flags(flag)
        try:
            K.wgrad(dy[:, :N], x, gw)
            torch.cuda.synchronize()
            ref = dy[:, :N].double().t() @ x.double()
            err = float((gw.double() - ref).abs().max() / ref.abs().max())
        except Exception as e:
            err = "failed: %s" % e
        print("flag", flag, "rows", rows, "N", N, "K", Kd, "rel err", err, flush=True)
L.lib().dtlr_debug_flags(0)

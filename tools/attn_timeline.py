"""Timeline of the single-pass tcgen05 attention kernel (mha_tc2_kernel): CTA (0,0,0) records clock() at fixed points of its MMA warp
and of two softmax warps (slot 0 / slot 3, lane quarter 2) -- dtlr_attn_debug_buffer.  Also prints the CUDA-graph time per layer of the
tcgen05 kernel and of the mma.sync flash kernel.  python tools/attn_timeline.py"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import _lib, ops  # noqa: E402
from gemm_probe_util import timeit  # noqa: E402

B, Q, heads, d = 64, 900, 8, 256
dt = torch.float16
qk = torch.randn(B * Q, 2 * d, device="cuda").to(dt)
v = torch.randn(B * Q, d, device="cuda").to(dt)
_lib.set_flavor(dt)
lib = _lib.lib()
for impl, flags, name in (("flash", 0, "mma.sync flash"), ("tc", 0, "tcgen05 single-pass")):
    ops.ATTN_IMPL = impl
    lib.dtlr_debug_flags(flags)
    us = timeit(lambda i: ops.mha_self_attention(qk, d, v, None, B, Q, heads, 32), iters=10)
    lib.dtlr_debug_flags(0)
    print("%s: %.1f us per layer" % (name, us), flush=True)
ops.ATTN_IMPL = "tc"
buf = torch.zeros(4 * 16 * 16, dtype=torch.int32, device="cuda")
lib.dtlr_attn_debug_buffer(ctypes.c_void_p(buf.data_ptr()))
ops.mha_self_attention(qk, d, v, None, B, Q, heads, 32)
torch.cuda.synchronize()
lib.dtlr_attn_debug_buffer(ctypes.c_void_p(0))
t = buf.cpu().numpy().astype(np.int64).reshape(4, 16, 16) & 0xFFFFFFFF
t0 = int(t[0, 0, 0])


def rel(x):
    return (int(x) - t0) & 0xFFFFFFFF if x else -1


print("MMA warp, S issue per chunk c: per slot (s_empty seen, issued) x 4")
for c in range(16):
    if t[0, c, 0]:
        print("  c %2d " % c + " ".join("%7d" % rel(t[0, c, s]) for s in range(8)))
print("MMA warp, P.V issue per chunk c (for chunk c-1): per slot (p_full seen, issued) x 4")
for c in range(16):
    if t[1, c, 0]:
        print("  c %2d " % c + " ".join("%7d" % rel(t[1, c, s]) for s in range(8)))
for role, name in ((2, "softmax warp of slot 0"), (3, "softmax warp of slot 3")):
    print("%s per chunk: start | s_full seen | tmem_ld done | max/exp/sum done | p_empty seen | O rescaled | P stored + fenced | arrived" % name)
    for c in range(16):
        if t[role, c, 0]:
            print("  c %2d " % c + " ".join("%7d" % rel(t[role, c, s]) for s in range(8)))

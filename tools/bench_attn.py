"""Decoder self-attention at BASELINE config-2 size (B=64, Q=900, 8 heads x 32): the mma.sync flash kernel (default), the
single-pass tcgen05 kernel (DTLR_ATTN=tc) and the two-pass tcgen05 kernel (dtlr_debug_flags(256)).  CUDA-graph timing."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import ops, _lib
from gemm_probe_util import timeit

B, Q, heads, d = 64, 900, 8, 256
qk = torch.randn(B * Q, 2 * d, device="cuda").bfloat16()
v = torch.randn(B * Q, d, device="cuda").bfloat16()
fl = 4.0 * B * heads * Q * Q * 32
for name, impl, flags in (("mma.sync flash (default)", "flash", 0), ("mma.sync flash, QK of the next key block software-pipelined", "flash", 16384),
                          ("tcgen05 single-pass", "tc", 0), ("tcgen05 two-pass", "tc", 256)):
    ops.ATTN_IMPL = impl
    _lib.lib().dtlr_debug_flags(flags)
    us = timeit(lambda i: ops.mha_self_attention(qk, d, v, None, B, Q, heads, 32), iters=10)
    _lib.lib().dtlr_debug_flags(0)
    print(json.dumps({"kernel": name, "us": round(us, 1), "TFLOPs": round(fl / us / 1e6, 1)}), flush=True)

# work partition of the flash kernel: warps per CTA x query splits per (image, head)  (0, 0 = automatic)
ops.ATTN_IMPL = "flash"
for W, S in ((16, 4), (0, 0), (19, 3), (20, 3), (15, 4), (19, 1), (20, 1), (16, 2), (18, 2), (12, 5), (19, 2), (15, 2)):
    _lib.lib().dtlr_attn_config(W, S)
    us = timeit(lambda i: ops.mha_self_attention(qk, d, v, None, B, Q, heads, 32), iters=10)
    print(json.dumps({"kernel": "flash warps=%d splits=%d" % (W, S), "us": round(us, 1), "TFLOPs": round(fl / us / 1e6, 1)}), flush=True)
_lib.lib().dtlr_attn_config(0, 0)

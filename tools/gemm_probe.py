"""Tuning probe: time dtlr_gemm with parts of the kernel disabled (dtlr_debug_flags) to see what bounds a tile."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import ops, _lib

def t(M, N, K, flags, iters=20):
    _lib.lib().dtlr_debug_flags(flags)
    a = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(4)]
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(4)]
    for i in range(4): ops.gemm(a[i % 4], w, bias, out=out[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): ops.gemm(a[i % 4], w, bias, out=out[i % 4])
    e1.record(); torch.cuda.synchronize()
    _lib.lib().dtlr_debug_flags(0)
    return e0.elapsed_time(e1) * 1000 / iters

for (M, N, K) in [(58368, 256, 256), (58368, 256, 2048), (58368, 2048, 256), (163840, 256, 64), (57600, 512, 256)]:
    print(M, N, K, {name: round(t(M, N, K, f), 1) for name, f in
                    [("full_bn256", 0), ("full_bn128", 8), ("no_store", 1), ("no_mma", 2), ("no_epilogue", 5), ("loads_only", 7)]})

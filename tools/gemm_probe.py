"""Tuning probe: time dtlr_gemm with parts of the kernel disabled (dtlr_debug_flags) to see what bounds a tile, next to
plain device copies of the same byte counts (what the memory system gives a kernel of this size)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import ops, _lib


def timeit(fn, iters=20):
    """GPU time per call of `iters` back-to-back launches replayed from a CUDA graph (no host launch gaps)."""
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(2):
            fn(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / iters


def t(M, N, K, flags, nbuf=6):
    _lib.lib().dtlr_debug_flags(flags)
    a = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(nbuf)]
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
    us = timeit(lambda i: ops.gemm(a[i % nbuf], w, bias, out=out[i % nbuf]))
    _lib.lib().dtlr_debug_flags(0)
    return us


def copy_ref(M, N, K, nbuf=6):
    """device copy moving the same bytes as the GEMM reads + writes (A in, C out)"""
    a = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(nbuf)]
    out = [torch.empty(M, K, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
    us = timeit(lambda i: out[i % nbuf].copy_(a[i % nbuf]))
    return us, 2 * M * K * 2 / us / 1e3


for (M, N, K) in [(58368, 256, 256), (58368, 2048, 256), (57600, 512, 256), (163840, 256, 64), (58368, 256, 2048)]:
    res = {name: round(t(M, N, K, f), 1) for name, f in
           [("ws", 0), ("ws_no_store", 1), ("ws_no_mma", 2), ("ws_no_epilogue", 4), ("ws_loads_only", 6),
            ("tile", 32), ("tile_no_store", 33), ("tile_no_mma", 34), ("tile_loads_only", 39)]}
    cu, gbs = copy_ref(M, N, K)
    print(M, N, K, res, "copy %dx%d bf16: %.1f us (%.0f GB/s)" % (M, K, cu, gbs), flush=True)


def ffn_probe(M=58368, hid=2048, nbuf=4):
    x = [torch.randn(M, 256, device="cuda").bfloat16() for _ in range(nbuf)]
    w1 = (torch.randn(hid, 256, device="cuda") / 16).bfloat16()
    w2 = (torch.randn(256, hid, device="cuda") / hid ** 0.5).bfloat16()
    b1, b2 = torch.randn(hid, device="cuda"), torch.randn(256, device="cuda")
    gm, bt = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
    fused = timeit(lambda i: ops.ffn_ln(x[i % nbuf], w1, b1, w2, b2, gm, bt), iters=10)
    ops.FFN_FUSED = False
    unf = timeit(lambda i: ops.ffn_ln(x[i % nbuf], w1, b1, w2, b2, gm, bt), iters=10)
    ops.FFN_FUSED = True
    fl = 4.0 * M * 256 * hid
    print("ffn M=%d hid=%d: fused %.1f us (%.0f TFLOP/s), linear1 + linear2/LN %.1f us" % (M, hid, fused, fl / fused / 1e6, unf), flush=True)


ffn_probe()

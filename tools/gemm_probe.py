"""Tuning probe: time dtlr_gemm with parts of the kernel disabled (dtlr_debug_flags) to see what bounds a tile, next to
plain device copies of the same byte counts (what the memory system gives a kernel of this size)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import ops, _lib


from gemm_probe_util import timeit  # noqa: E402


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


import sys as _s
for (M, N, K) in [] if (len(_s.argv) > 1 and _s.argv[1] == 'ffn') else [(58368, 256, 256), (58368, 2048, 256), (57600, 512, 256), (163840, 256, 64), (58368, 256, 2048)]:
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
    ops.FFN_FUSED = True
    fused = timeit(lambda i: ops.ffn_ln(x[i % nbuf], w1, b1, w2, b2, gm, bt), iters=10)
    probes = {}
    for name, f in [("cta_pairs_multicast", 4096), ("no_E1", 256), ("no_G1", 512), ("no_G2", 1024), ("no_final", 2048), ("no_mma", 1536), ("no_mma_no_E1", 1792), ("loads_only", 3840)]:
        _lib.lib().dtlr_debug_flags(f)
        probes[name] = round(timeit(lambda i: ops.ffn_ln(x[i % nbuf], w1, b1, w2, b2, gm, bt), iters=10), 1)
    _lib.lib().dtlr_debug_flags(0)
    print("ffn probes", probes, flush=True)
    ops.FFN_FUSED = False
    unf = timeit(lambda i: ops.ffn_ln(x[i % nbuf], w1, b1, w2, b2, gm, bt), iters=10)
    fl = 4.0 * M * 256 * hid
    print("ffn M=%d hid=%d: fused %.1f us (%.0f TFLOP/s), linear1 + linear2/LN %.1f us" % (M, hid, fused, fl / fused / 1e6, unf), flush=True)


ffn_probe()


def ln_probe(M=58368, nbuf=6):
    a = [torch.randn(M, 256, device="cuda").bfloat16() for _ in range(nbuf)]
    r = [torch.randn(M, 256, device="cuda").bfloat16() for _ in range(nbuf)]
    w = (torch.randn(256, 256, device="cuda") / 16).bfloat16()
    b = torch.randn(256, device="cuda")
    gm, bt = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
    ops.LN_FUSE_WS = True
    fused = timeit(lambda i: ops.linear_ln(a[i % nbuf], w, b, r[i % nbuf], gm, bt))
    fused_nores = timeit(lambda i: ops.linear_ln(a[i % nbuf], w, b, None, gm, bt))
    ops.LN_FUSE_WS = False
    unf = timeit(lambda i: ops.linear_ln(a[i % nbuf], w, b, r[i % nbuf], gm, bt))
    ops.LN_FUSE_WS = True
    print("linear(256->256)+res+LN M=%d: fused %.1f us (no residual %.1f), gemm + add_layernorm %.1f us" % (M, fused, fused_nores, unf), flush=True)


ln_probe()


def ln2048_probe(M=58368, nbuf=4):
    a = [torch.randn(M, 2048, device="cuda").bfloat16() for _ in range(nbuf)]
    r = [torch.randn(M, 256, device="cuda").bfloat16() for _ in range(nbuf)]
    w = (torch.randn(256, 2048, device="cuda") / 45).bfloat16()
    b = torch.randn(256, device="cuda")
    gm, bt = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
    plain = timeit(lambda i: ops.gemm(a[i % nbuf], w, b, residual=r[i % nbuf]), iters=10)
    ln = timeit(lambda i: ops.gemm_ln(a[i % nbuf], w, b, r[i % nbuf], gm, bt), iters=10)
    ln2 = timeit(lambda i: ops.gemm_ln(a[i % nbuf], w, b, r[i % nbuf], gm, bt, r[(i + 1) % nbuf]), iters=10)
    print("linear2 (2048->256) M=%d: gemm+res %.1f us, gemm_ln %.1f us, gemm_ln+add2 %.1f us" % (M, plain, ln, ln2), flush=True)


ln2048_probe()

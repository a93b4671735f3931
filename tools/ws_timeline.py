"""Timeline of the weight-stationary tcgen05 GEMM (gemm_ws_tcgen05_kernel): CTA 0 records clock() at fixed points of its producer
thread, MMA warp and epilogue warp 2 (dtlr_gemm_debug_buffer).  python tools/ws_timeline.py [M N K [residual 0/1]]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import _lib, ops  # noqa: E402

M, N, K = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (58368, 256, 256)
res = len(sys.argv) > 4 and sys.argv[4] == "1"
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dt = torch.float16
g = torch.Generator(device="cuda").manual_seed(0)
a = [torch.randn(M, K, device="cuda", generator=g).to(dt) for _ in range(3)]
w = (torch.randn(N, K, device="cuda", generator=g) / 16).to(dt)
b = torch.randn(N, device="cuda", generator=g) * 0.1
r = torch.randn(M, N, device="cuda", generator=g).to(dt) if res else None
_lib.set_flavor(dt)
lib = _lib.lib()
lib.dtlr_debug_flags(flags)
buf = torch.zeros(4 * 64 * 16, dtype=torch.int32, device="cuda")
for i in range(4):
    ops.gemm(a[i % 3], w, b, residual=r)
torch.cuda.synchronize()
# graph-timed reference number
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for i in range(12):
        ops.gemm(a[i % 3], w, b, residual=r)
gr.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
print("M %d N %d K %d residual %s flags %d: %.1f us per call in a graph of 12" % (M, N, K, res, flags, e0.elapsed_time(e1) / 12 * 1e3))
if os.environ.get("WS_TIME_ONLY"):
    sys.exit(0)
lib.dtlr_gemm_debug_buffer(ctypes.c_void_p(buf.data_ptr()))
ops.gemm(a[1], w, b, residual=r)
ops.gemm(a[2], w, b, residual=r)      # the second launch overwrites: it is the one that overlapped a predecessor
torch.cuda.synchronize()
lib.dtlr_gemm_debug_buffer(ctypes.c_void_p(0))
t = buf.cpu().numpy().astype(np.int64).reshape(4, 64, 16) & 0xFFFFFFFF
t0 = int(t[3, 0, 1])      # MMA warp's first stamp at kernel entry


def rel(x):
    return (int(x) - t0) & 0xFFFFFFFF if x else -1


print("kernel entry (warp 1) 0; after pdl_wait %d; epilogue warp 2 done %d" % (rel(t[3, 1, 0]), rel(t[3, 2, 0])))
print("producer per tile: start | empty kb0..3")
for i in range(8):
    if t[0, i, 0]:
        print("  tile %d " % i + " ".join("%7d" % rel(t[0, i, s]) for s in range(5)))
print("MMA per tile: start | tmem_empty | kb0 full, issued | kb1 | kb2 | kb3")
for i in range(8):
    if t[1, i, 0]:
        print("  tile %d " % i + " ".join("%7d" % rel(t[1, i, s]) for s in range(10)))
print("epilogue warp 2 per tile: start | blk0: tmem_full/before ld, after ld, math, wait_read, staged+stored | blk1 ...")
for i in range(8):
    if t[2, i, 0]:
        print("  tile %d " % i + " ".join("%7d" % rel(t[2, i, s]) for s in range(11)))

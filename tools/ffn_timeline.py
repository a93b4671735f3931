"""Timeline of the stream-K FFN kernel on CTA pairs: CTA 0 (the leader of pair 0) records clock() at fixed points of its TMA-producer
thread, MMA-issuer warp and epilogue warp 2 (dtlr_ffn_debug_buffer).  Prints, per (row tile, hidden chunk) unit, where each role
waits -- cycles of the SM clock.  python tools/ffn_timeline.py [flags]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import _lib, ops  # noqa: E402

M, hid = 58368, 2048
dt = torch.float16
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 1073741824
g = torch.Generator(device="cuda").manual_seed(0)
x = [torch.randn(M, 256, device="cuda", generator=g).to(dt) for _ in range(3)]
w1 = (torch.randn(hid, 256, device="cuda", generator=g) / 16).to(dt)
w2 = (torch.randn(256, hid, device="cuda", generator=g) / 45).to(dt)
b1 = torch.randn(hid, device="cuda", generator=g) * 0.1
b2 = torch.randn(256, device="cuda", generator=g) * 0.1
gm = torch.ones(256, device="cuda")
bt = torch.zeros(256, device="cuda")
_lib.set_flavor(dt)
lib = _lib.lib()
lib.dtlr_debug_flags(flags)
buf = torch.zeros(3 * 64 * 16, dtype=torch.int32, device="cuda")
for i in range(4):
    ops.ffn_ln(x[i % 3], w1, b1, w2, b2, gm, bt)
torch.cuda.synchronize()
lib.dtlr_ffn_debug_buffer(ctypes.c_void_p(buf.data_ptr()))
ops.ffn_ln(x[1], w1, b1, w2, b2, gm, bt)
torch.cuda.synchronize()
lib.dtlr_ffn_debug_buffer(ctypes.c_void_p(0))
t = buf.cpu().numpy().astype(np.int64).reshape(3, 64, 16) & 0xFFFFFFFF
prod, mma, epi = t[0], t[1], t[2]
t0 = int(mma[0, 0])


def rel(a):
    return (int(a) - t0) & 0xFFFFFFFF if a else -1


print("flags", flags, "plan", lib.dtlr_ffn_plan(M, hid))
print("MMA warp per iteration v: start | x_full | G1 st0: w_full, issued | st1: w_full, issued | h_full(v-1) | y_free | G2 kb0: w_full, issued | kb1: w_full, issued")
for v in range(0, 52):
    if mma[v, 0] == 0:
        break
    print("v %2d  " % v + " ".join("%7d" % rel(mma[v, s]) for s in range(12)), " | iteration %5d clk" % ((int(mma[v + 1, 0]) - int(mma[v, 0])) & 0xFFFFFFFF if mma[v + 1, 0] else -1))
print("epilogue warp 2 per unit v: start | hacc_full | tmem_ld | math + bar | tmem_st | arrive")
for v in range(0, 52):
    if epi[v, 0] == 0:
        break
    print("v %2d  " % v + " ".join("%7d" % rel(epi[v, s]) for s in range(6)),
          " | hacc commit -> seen %5d, E1 %5d, arrive -> MMA sees h_full %5d" % (
              (int(epi[v, 1]) - int(mma[v, 5])) & 0xFFFFFFFF if mma[v, 5] else -1, int(epi[v, 5]) - int(epi[v, 1]),
              (int(mma[v + 1, 6]) - int(epi[v, 5])) & 0xFFFFFFFF if mma[v + 1, 6] else -1))
print("final epilogues (item t): start | y_full | first pass | second pass | stores issued + read")
for tt in range(0, 6):
    if epi[56 + tt, 0] == 0:
        continue
    print("t %d  " % tt + " ".join("%7d" % rel(epi[56 + tt, s]) for s in range(5)))
print("producer per iteration v: start | W1 st0 w_empty | st1 | - | - | W2 q0 w_empty | q1 | - | - | before X prefetch | after")
for v in range(0, 52):
    if prod[v, 0] == 0:
        break
    print("v %2d  " % v + " ".join("%7d" % rel(prod[v, s]) for s in range(11)))

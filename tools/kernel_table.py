"""Per-call timing table of one eager forward (B=64, 3x40x1024, bf16): every C-ABI call is bracketed by a CUDA-event pair on
its stream (no profiler), grouped by entry point and problem size.  Shows where a step's time goes and each GEMM shape's
achieved TFLOP/s and GB/s.   python tools/kernel_table.py [batch]"""
import collections
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import _lib, config, dino, synth  # noqa: E402

RECORDS = []


class Proxy:
    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        fn = getattr(self._real, name)
        if not name.startswith("dtlr_") or name in ("dtlr_last_error", "dtlr_version", "dtlr_debug_flags"):
            return fn

        def wrapped(*args):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            ints = tuple(a if isinstance(a, int) else (a.value if isinstance(a, (ctypes.c_longlong, ctypes.c_int)) else None) for a in args)
            RECORDS.append((name, ints, e0, e1))
            return rc
        return wrapped


def key_of(name, a):
    if name == "dtlr_gemm":
        return "gemm M=%d N=%d K=%d out=%s relu=%d res=%d" % (a[9], a[10], a[11], "f32" if a[13] == 0 else "bf16", a[14], 0), (a[9], a[10], a[11], a[13])
    if name == "dtlr_gemm_ln":
        return "gemm_ln M=%d N=256 K=%d" % (a[-3], a[-2]), (a[-3], 256, a[-2], 1)
    if name == "dtlr_conv2d_nhwc":
        B, H, W, C, Cout, KH, KW = a[5:12]
        return "conv3x3 M=%d N=%d K=%d" % (B * H * W, Cout, KH * KW * C), (B * H * W, Cout, KH * KW * C, 1)
    if name in ("dtlr_msda_forward_fused",):
        return "msda_fused Lq=%d ref=%d" % (a[16], a[8]), None
    if name == "dtlr_mha_self_attention":
        return "mha B=%d Q=%d" % (a[8], a[9]), None
    return name[5:], None


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = torch.device("cuda", 0)
    model, _, _ = dino.build_dino(config.latin_ctc_args())
    synth.load_synth_weights(model, seed=0)
    model = model.to(dev).eval()
    model.compute_dtype = torch.bfloat16
    model.engine_outputs = "all"
    model.use_cuda_graph = False
    x = synth.synth_images(B, 40, 1024, seed=1).to(dev)
    with torch.no_grad():
        for _ in range(2):
            model(x)
        torch.cuda.synchronize()
        real = _lib.lib()
        _lib._lib = Proxy(real)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        model(x)
        t1.record()
        torch.cuda.synchronize()
        _lib._lib = real
    agg = collections.OrderedDict()
    for name, ints, e0, e1 in RECORDS:
        k, shape = key_of(name, ints)
        d = agg.setdefault(k, {"n": 0, "us": 0.0, "shape": shape})
        d["n"] += 1
        d["us"] += e0.elapsed_time(e1) * 1e3
    total = sum(d["us"] for d in agg.values())
    print("# eager forward B=%d: wall %.2f ms, sum of bracketed calls %.2f ms, %d calls" % (B, t0.elapsed_time(t1), total / 1e3, len(RECORDS)))
    print("# %-52s %5s %10s %9s %6s %9s %8s" % ("call", "n", "total us", "us/call", "%", "TFLOP/s", "GB/s"))
    for k, d in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        tf = gb = ""
        if d["shape"]:
            M, N, K, odt = d["shape"]
            per = d["us"] / d["n"]
            tf = "%.0f" % (2.0 * M * N * K / per / 1e6)
            gb = "%.0f" % ((M * K * 2 + N * K * 2 + M * N * (4 if odt == 0 else 2)) / per / 1e3)
        print("  %-52s %5d %10.1f %9.1f %6.1f %9s %8s" % (k, d["n"], d["us"], d["us"] / d["n"], 100 * d["us"] / total, tf, gb))


if __name__ == "__main__":
    main()

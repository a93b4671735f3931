"""Split-precision mode (model.split_precision) at the bench shape: ms per step under CUDA-graph replay, and a kernel table of one
eager step (torch.profiler) -- where the tensor-core parity mode spends its time.  usage: python tools/bench_split.py [table | shapes]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dtlr_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
model = bench.build_ours(dev, torch.float32)
model.split_precision = True
x = synth.synth_images(bench.BATCH_PER_GPU, bench.IMG_H, bench.IMG_W, seed=100).to(dev)
with torch.no_grad():
    for _ in range(3):
        model(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        model(x)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("split mode B=%d: %.2f ms per step, %.0f img/s (implicit conv %s)" % (x.shape[0], ms, x.shape[0] / ms * 1e3, os.environ.get("DTLR_SPLIT_CONV_IMPLICIT", "1")))
if len(sys.argv) > 1 and sys.argv[1] == "table":
    model.use_cuda_graph = False
    from torch.profiler import ProfilerActivity, profile
    with torch.no_grad():
        model(x)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            model(x)
            torch.cuda.synchronize()
    rows = sorted(((e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0), key=lambda r: -r[2])
    tot = sum(r[2] for r in rows)
    print("eager step: %.2f ms of kernels" % (tot / 1e3))
    for k, n, t in rows[:28]:
        print("%7.1f us %5.1f %% x%-4d %s" % (t, 100 * t / tot, n, k[:150]))

if len(sys.argv) > 1 and sys.argv[1] == "shapes":
    # per-call CUDA-event timing of every contraction of one eager step, aggregated by shape
    import collections
    from dtlr_b200 import ops
    model.use_cuda_graph = False
    rec = []
    og, oc = ops.gemm, ops.conv2d_nhwc

    def timed(kind, fn, shape, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        rec.append((kind, shape, e0, e1))
        return r

    def gemm(a, w, *args, **kw):
        return timed("gemm", og, (a.shape[0], w.shape[0], a.shape[1], str(a.dtype)[6:], str(kw.get("out_dtype"))[-8:], "res" if kw.get("residual") is not None else ""), a, w, *args, **kw)

    def conv(x, w, bias, B, H, W, C, k, pad, **kw):
        return timed("conv", oc, (x.shape[0], w.shape[0], w.shape[1], "s%d" % kw.get("stride", 1), str(kw.get("out_dtype"))[-8:], ""), x, w, bias, B, H, W, C, k, pad, **kw)

    ops.gemm, ops.conv2d_nhwc = gemm, conv
    with torch.no_grad():
        model(x)
        rec.clear()
        model(x)
        torch.cuda.synchronize()
    ops.gemm, ops.conv2d_nhwc = og, oc
    agg = collections.OrderedDict()
    for kind, shape, e0, e1 in rec:
        t = agg.setdefault((kind,) + shape, [0, 0.0])
        t[0] += 1
        t[1] += e0.elapsed_time(e1) * 1e3
    tot = sum(v[1] for v in agg.values())
    print("contractions of one eager step (incl. the split pass of fp32 A operands): %.2f ms in %d calls" % (tot / 1e3, len(rec)))
    for key, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        kind, M, N, K = key[:4]
        kl = K // 3 if key[4] in ("float16", "s1", "s2") else K
        print("%8.1f us %5.1f %% x%-3d %7.1f us each  %s M=%d N=%d K=%d %s  %.0f TFLOP/s (algorithmic)" % (
            t, 100 * t / tot, n, t / n, kind, M, N, K, " ".join(str(v) for v in key[4:]), 2.0 * M * N * kl * n / t / 1e6))

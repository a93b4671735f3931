"""Measurements for the SURVEY 8f rows either side of the forward (one JSON line each):
  1. dtlr_preprocess_u8 (GPU input stage) at B=64, 40x1024 grayscale: CUDA-event time, algorithmic GB/s against the measured copy peak;
  2. LineEvaluator over real-resolution lines (reference evaluation geometry: height 94, widths up to 1333, datasets/IAM.py:225-229)
     in (width, height) buckets of up to 16 lines vs the reference's loop shape (one image per forward, evaluation.py:494-499) on the same
     model -- bf16 engine, random weights, synthetic u8 lines; whole-loop wall time bracketed by synchronize.
     python tools/bench_eval.py [n_lines]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import config, dino, evaluation, ops, synth  # noqa: E402
from dtlr_b200.input import IMAGENET_MEAN, IMAGENET_STD, pack_u8  # noqa: E402

PEAK = 6467.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def bench_preprocess(B=64, H=40, W=1024):
    rng = np.random.default_rng(0)
    packed, off, sizes, ch = pack_u8([rng.integers(0, 256, (H, W), dtype=np.uint8) for _ in range(B)])
    dp = torch.from_numpy(packed).cuda()
    do = torch.from_numpy(off).cuda()
    dhw = torch.from_numpy(sizes).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ops.preprocess_u8(dp, do, dhw, ch, B, H, W, IMAGENET_MEAN, IMAGENET_STD)
    ts = []
    for _ in range(10):
        flush.zero_()                                   # evict the input / output lines from the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.preprocess_u8(dp, do, dhw, ch, B, H, W, IMAGENET_MEAN, IMAGENET_STD)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = float(np.median(ts))
    nbytes = B * H * W * 13 + packed.size
    print(json.dumps({"op": "dtlr_preprocess_u8", "B": B, "H": H, "W": W, "us": round(us, 2), "algorithmic_bytes": nbytes,
                      "GBps": round(nbytes / us / 1e3, 1), "frac_of_hbm_peak": round(nbytes / us / 1e3 / PEAK, 3),
                      "h2d_bytes_u8": int(packed.size), "h2d_bytes_fp32_reference_path": B * 3 * H * W * 4,
                      "l2": "256 MB flush write between timed launches"}), flush=True)


def bench_evaluator(n_lines=256):
    model, _, _ = dino.build_dino(config.latin_ctc_args())
    synth.load_synth_weights(model, seed=0)
    model = model.cuda().eval()
    model.compute_dtype = torch.bfloat16
    model.use_cuda_graph = True
    rng = np.random.default_rng(1)
    # original scans: heights ~ N(120, 30), widths ~ N(1750, 450); the evaluation transform (datasets/IAM.py:225-229:
    # RandomResize([800], max_size=1333)) maps them to width 1333 (long side capped) and heights 50..200
    from dtlr_b200.input import get_size_with_aspect_ratio
    sizes = []
    while len(sizes) < n_lines:
        h0, w0 = int(rng.normal(120, 30)), int(rng.normal(1750, 450))
        if h0 < 40 or w0 < 600:
            continue
        oh, ow = get_size_with_aspect_ratio((w0, h0), 800, 1333)
        if 48 <= oh <= 224:          # fewer than 900 encoder tokens makes topk(900) fail in the reference as well
            sizes.append((oh, ow))
    lines = [rng.integers(0, 256, s, dtype=np.uint8) for s in sizes]
    charset = [chr(0x21 + i) for i in range(166)]
    res = {}
    for name, bs in (("bucketed_16", 16), ("one_image_per_forward", 1)):
        ev = evaluation.LineEvaluator(model, charset, batch_size=bs, width_multiple=64, height_multiple=16)
        ev.predict(lines)                               # warm-up: captures one CUDA graph per (batch, size bucket) shape
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        preds = ev.predict(lines)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        bt = ev.batches(lines)
        res[name] = {"images_per_s": round(n_lines / dt, 1), "batches": len(bt),
                     "shapes": len({(len(b), -(-max(sizes[i][0] for i in b) // 16), -(-max(sizes[i][1] for i in b) // 64)) for b in bt})}
        assert len(preds) == n_lines
        if bs == 16:
            px = sum(h * w for h, w in sizes)
            padded = sum(len(b) * (-(-max(sizes[i][0] for i in b) // 16) * 16) * (-(-max(sizes[i][1] for i in b) // 64) * 64) for b in bt)
            res["padding_overhead_bucketed"] = round(padded / px - 1, 4)
    print(json.dumps({"op": "LineEvaluator.predict", "lines": n_lines, "mean_hw": [float(np.mean([s[0] for s in sizes])),
                      float(np.mean([s[1] for s in sizes]))], "dtype": "bf16", **res}), flush=True)


if __name__ == "__main__":
    bench_preprocess()
    bench_preprocess(32, 94, 1333)
    bench_evaluator(int(sys.argv[1]) if len(sys.argv) > 1 else 256)

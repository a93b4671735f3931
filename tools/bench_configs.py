"""Device-resident forward + decode tail on the other BASELINE configurations (parity-test cases, not bench lines):
  H: config/HWDB_full.py, 7356 classes, batch 32 of 40x1024 -- stresses the C-wide class heads (6 decoder layers + 2 encoder-side
     heads) and the decode tail (HBM bound, B*Q*C*4 bytes of logits read per call);
  A: config/Latin_CTC.py batch 64 for the decode tail at C = 166.
CUDA-graph replay for the forward, CUDA events, 10 timed iterations after 3 warm-ups.  One JSON line per measurement."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import config, dino, ops, synth  # noqa: E402

PEAK = 6467.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(name, args, B, outputs):
    model, _, _ = dino.build_dino(args)
    synth.load_synth_weights(model, seed=0)
    model = model.cuda().eval()
    model.compute_dtype = torch.bfloat16
    model.engine_outputs = outputs
    model.use_cuda_graph = True
    x = synth.synth_images(B, 40, 1024, seed=3).cuda()
    with torch.no_grad():
        ms = timed(lambda: model(x))
        out = model(x)
        C = out["pred_logits"].shape[-1]
        lg, bx = out["pred_logits"].float().contiguous(), out["pred_boxes"].float().contiguous()
        flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

        def dec():
            flush.zero_()
            return ops.ctc_decode(lg, bx, 0.003)
        ms_flush = timed(lambda: flush.zero_())
        ms_dec_vec = timed(dec) - ms_flush          # default: row-label kernel with 16-byte loads (rows 16-byte aligned)
        from dtlr_b200 import _lib
        _lib.lib().dtlr_debug_flags(32768)          # scalar-load kernel (A/B)
        ms_dec = timed(dec) - ms_flush
        _lib.lib().dtlr_debug_flags(0)
    nbytes = B * 900 * (C * 4 + 16 + 4)
    print(json.dumps({"config": name, "batch": B, "classes": C, "outputs": outputs, "dtype": "bf16", "forward_ms": round(ms, 3),
                      "images_per_s": round(B / ms * 1e3, 1), "decode_us_scalar_loads": round(ms_dec * 1e3, 1), "decode_us": round(ms_dec_vec * 1e3, 1),
                      
                      "decode_algorithmic_GBps": round(nbytes / ms_dec_vec / 1e6, 1), "decode_frac_of_hbm_peak": round(nbytes / ms_dec_vec / 1e6 / PEAK, 3),
                      "decode_l2": "512 MB flush write before every timed decode (its time subtracted)"}), flush=True)
    del model
    torch.cuda.empty_cache()


if __name__ == "__main__":
    run("H (HWDB_full, 7356 classes)", config.hwdb_args(), 32, "all")
    run("A (Latin_CTC, 166 classes)", config.latin_ctc_args(), 64, "all")

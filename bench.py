#!/usr/bin/env python
"""bench.py -- BASELINE.json headline metric: text-line images/sec, DINO forward, batch 64 of 3x40x1024 per GPU
(config/Latin_CTC.py: ResNet-50 + 6+6 deformable enc/dec, 900 queries, 166 classes), bf16, synthetic data.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run, one rank/GPU)
  python bench.py --impl reference ...                     (the reference path's CPU restatement on the host cores)

One "step" = one full forward (all reference outputs: 6 decoder layers of logits/boxes + interm outputs) of one batch.
Prints ONE JSON line (rank 0).  See DESIGN.md §measurement for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH_PER_GPU = 64
IMG_H, IMG_W = 40, 1024
import glob
import re

_UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}


def profile_metrics(kernel_regex, metrics, hint="", prefer=("r2_", "r1_")):
    """Read per-launch ncu metrics of one kernel from the committed summaries under profiles/ (written by tools/ncu_summary.py from
    an `ncu --set full` capture; newest round first).  Returns ({metric: value in base units}, file) or ({}, None)."""
    pat = re.compile(kernel_regex)
    files = []
    for pre in prefer:
        fs = sorted(glob.glob(os.path.join(ROOT, "profiles", pre + "*ncu*.txt")), reverse=True)
        files += [f for f in fs if hint and hint in os.path.basename(f)] + [f for f in fs if not (hint and hint in os.path.basename(f))]
    for f in files:
        block = None
        found = {}
        try:
            lines = open(f).read().splitlines()
        except OSError:
            continue
        for ln in lines:
            if ln.startswith("===="):
                if found:
                    break
                block = pat.search(ln) is not None
                continue
            if block:
                parts = ln.split()
                if len(parts) >= 2 and parts[0] in metrics:
                    try:
                        val = float(parts[1].replace(",", ""))
                    except ValueError:
                        continue
                    found[parts[0]] = val * _UNIT.get(parts[2] if len(parts) > 2 else "", 1.0)
        if found:
            return found, os.path.relpath(f, ROOT)
    return {}, None


def profile_traffic(kernel_regex, hint=""):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the kernel, from the newest committed ncu summary"""
    m, f = profile_metrics(kernel_regex, ("dram__bytes_read.sum", "dram__bytes_write.sum"), hint)
    if len(m) == 2:
        return m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"], f
    return None, None
WORKLOAD = "IAM English config/Latin_CTC.py: ResNet-50 + 6+6 deformable enc/dec, 900 queries, 166 classes, 64x3x40x1024 per GPU, forward"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed regions (B200_PROFILING.md).  The sampler process is started
    before the warm-up (nvidia-smi needs ~0.5 s to produce its first line) and every line is time-stamped on arrival; only lines
    that arrived inside a window opened by begin() / closed by end() are reported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines, self.windows, self._t0 = gpu_index, None, [], [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def begin(self):
        self._t0 = time.perf_counter()

    def end(self):
        if self._t0 is not None:
            self.windows.append((self._t0, time.perf_counter()))
            self._t0 = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, ln in self.lines:
            if not any(a <= ts <= b + 0.02 for a, b in self.windows):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "window": "device-resident timed region + end-to-end timed regions (nvidia-smi -lms 20)"}


def gemm_hbm_view(gemm_bytes, gemm_ms, n_prof, hbm_peak_gbs):
    """HBM reading of the dtlr_gemm family: sum of algorithmic bytes / sum of event-timed durations against the copy peak."""
    gbs = gemm_bytes / (gemm_ms * 1e-3) / 1e9 if gemm_ms > 0 else 0.0
    return {"achieved": round(gbs, 1), "peak": hbm_peak_gbs, "unit": "GB/s", "frac": round(gbs / hbm_peak_gbs, 4),
            "algorithmic_bytes_per_step": gemm_bytes / max(1, n_prof),
            "what": "sum over the same launches of A + W + output (+ residual) bytes / sum of durations, against the measured copy peak"}


def msda_binding_view(us_per_launch, sms, sm_mhz, wavefronts, source):
    """the resource that binds the deformable-attention core: shared-memory wavefronts, one 128-byte wavefront per clock per SM
    (wavefronts = ncu l1tex__data_pipe_lsu_wavefronts_mem_shared.sum of one launch, read from the committed summary `source`)"""
    floor_us = wavefronts / sms / sm_mhz
    return {"us_per_launch": round(us_per_launch, 1), "binding_floor_us": round(floor_us, 1),
            "binding_frac": round(floor_us / us_per_launch, 3), "shared_wavefronts_per_launch": wavefronts, "wavefronts_from": source}


def build_ours(device, dtype):
    from dtlr_b200 import config, dino, synth
    model, criterion, post = dino.build_dino(config.latin_ctc_args())
    synth.load_synth_weights(model, seed=0)
    model = model.to(device).eval()
    model.compute_dtype = dtype
    model.engine_outputs = "all"
    model.use_cuda_graph = os.environ.get("DTLR_NO_GRAPH", "0") != "1"
    return model


def cpu_baseline_run(batch, iters, warmup=1):
    """the oracle (CPU restatement of the reference path: torch CPU ops + C MSDA core) on the host cores."""
    import json as _json
    from dtlr_b200 import synth
    from oracle import dino_ref
    shapes = _json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    sd = synth.synth_state_dict(shapes, seed=0)
    cfg = dino_ref.default_cfg(num_queries=900)
    x = synth.synth_images(batch, IMG_H, IMG_W, seed=0)
    # all the host threads the restatement can use: its small per-layer ops stop scaling past ~16-32 threads (measured on the
    # 128-core B200 host: 8 threads 5.7 img/s, 16: 6.3, 32: 6.2, 64: 3.7, 128: 0.38), so the baseline runs at its best setting;
    # both numbers are reported: `cores` = threads used, `host_cores` = os.cpu_count()
    torch.set_num_threads(min(os.cpu_count() or 1, 16))
    for _ in range(warmup):
        dino_ref.dino_forward(sd, cfg, x)
    t0 = time.perf_counter()
    for _ in range(iters):
        dino_ref.dino_forward(sd, cfg, x)
    dt = time.perf_counter() - t0
    return batch * iters / dt, dt / iters


def run_reference(args, rank):
    if rank != 0:
        return
    batch = 8
    ips, sec = cpu_baseline_run(batch, max(1, args.steps), max(1, min(args.warmup, 1)))
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": "text-line images/sec (DINO forward)", "value": round(ips, 3), "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "step": "bounded sample: %d images per step on the host CPU" % batch},
            "cpu_baseline": {"value": round(ips, 3), "unit": "images/s", "cores": cores, "host_cores": os.cpu_count(), "kind": "port",
                             "why_port": "the reference is a Python source tree with no setup.py / pyproject (not pip-installable) and cannot travel to the GPU box; its CPU restatement, pinned to reference-generated fixtures, is timed",
                             "sample": "%d steps of %d images (3x40x1024), oracle/dino_ref.py + oracle/msda_ref.c, fp32" % (args.steps, batch)},
            "e2e": {"value": round(ips, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


TRAIN_BATCH_PER_GPU = 32


def train_step_leg(device, rank, world, local, steps=5, warmup=3, impl="native"):
    """BASELINE config 5: IAM fine-tune step = forward(samples, targets) + loss_CTC + backward + clip_grad_norm(0.01) + AdamW, 32 lines
    per GPU (reference engine.py:172-274, finetuning.py:211-231).  Timed on the device, max over ranks.
    impl "native": dtlr_b200.train_engine.TrainEngine -- transformer forward AND backward, loss and optimizer on libdtlr_b200 kernels
    (bf16 tcgen05 operands, fp32 accumulation / gradient arena), ResNet front under torch autograd; N > 1: the flat gradient arena is
    all-reduced over NCCL in three segments overlapped with the backward.
    impl "torch": the module path under torch autograd (cuBLAS / cuDNN, TF32) with DistributedDataParallel -- the round-1 step, kept as
    the A/B."""
    from dtlr_b200 import config, dino, dist_util, synth
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    model, crit, _ = dino.build_dino(config.latin_ctc_args())
    synth.load_synth_weights(model, seed=0)
    model = model.to(device).train()
    B = TRAIN_BATCH_PER_GPU
    x = synth.synth_images(B, IMG_H, IMG_W, seed=300 + rank).to(device)
    tg = [{k: v.to(device) for k, v in t.items()} for t in synth.synth_targets(B, 166, seed=300 + rank)]
    if impl == "native":
        from dtlr_b200 import train_engine
        eng = train_engine.TrainEngine(model, lr=1e-5, lr_backbone=1e-10, weight_decay=1e-4, max_norm=0.01, dtype=torch.bfloat16,
                                       world_size=world)
        n_grad = eng.n_live

        def step():
            return eng.step(x, tg)
        what = ("forward(samples, targets) + loss_CTC + backward + clip + AdamW on dtlr kernels in both directions: ResNet-50 layer2-4 + "
                "input_proj / GroupNorm, encoder, decoder (flash self-attention forward + backward with the DN mask), heads -- tcgen05 GEMM / "
                "implicit-GEMM conv forward, dgrad on the same kernels, MN-major tcgen05 wgrad, LayerNorm / GroupNorm / ReLU / MSDA / CTC "
                "backward kernels, fused clip + AdamW over flat fp32 arenas; bf16 operands, fp32 accumulation and gradient stream; torch "
                "only for index bookkeeping (masks, DN query table, embedding-table gradients)")
        coll = ("%d async NCCL all-reduces of the flat fp32 gradient arena (%d gradients, %.0f MB) issued as each segment's backward "
                "completes (decoder | encoder | front)" % (len(eng.chunks), n_grad, n_grad * 4 / 1e6))
    else:
        net = model
        if world > 1:
            from torch.nn.parallel import DistributedDataParallel as DDP
            net = DDP(model, device_ids=[local], find_unused_parameters=True, gradient_as_bucket_view=True, bucket_cap_mb=64)   # (finetuning.py:211-215)
        params = [p for p in net.parameters() if p.requires_grad]
        opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=1e-4)
        n_grad = sum(p.numel() for p in params)

        def step():
            opt.zero_grad(set_to_none=True)
            out = net(x, tg)
            loss = crit.loss_CTC(out, tg, None, None)["loss_CTC"]
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 0.01)
            opt.step()
            return loss
        what = ("forward(samples, targets) + loss_CTC + backward + clip + AdamW; dtlr kernels: deformable attention fwd/bwd, fused CTC "
                "loss fwd/bwd; other layers torch autograd over cuBLAS/cuDNN (TF32)")
        coll = "DDP bucketed NCCL all-reduce of %d fp32 gradients (%.0f MB) per step, overlapped with backward" % (n_grad, n_grad * 4 / 1e6)

    for _ in range(warmup):
        step()
    dist_util.barrier(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    dist_util.barrier(device)
    ms = dist_util.max_over_ranks(e0.elapsed_time(e1), device) / steps
    res = {"value": round(world * B / ms * 1e3, 1), "unit": "images/s", "ms_per_step": round(ms, 2), "batch_per_gpu": B,
           "global_batch": B * world, "loss": round(float(loss), 4), "steps": steps, "warmup": warmup, "impl": impl,
           "collective": coll if world > 1 else "none (N = 1)", "what": what}
    del model
    torch.cuda.empty_cache()
    return res


def gpu_reference_leg(device, batch, iters=3):
    """The reference-shaped torch path on the SAME GPU: oracle/dino_ref.py run on CUDA tensors (torch fp32 ops over cuBLAS / cuDNN, TF32
    off) with the reference's OWN deformable-attention CUDA kernel recompiled for sm_100a (oracle/_ref).  A like-for-like baseline for the
    fused engine, reported beside the CPU arm; test infrastructure, never on the product path."""
    import json as _json
    from dtlr_b200 import synth
    from oracle import dino_ref
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    shapes = _json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    sd = {k: v.to(device) for k, v in synth.synth_state_dict(shapes, seed=0).items()}
    cfg = dino_ref.default_cfg(num_queries=900)
    x = synth.synth_images(batch, IMG_H, IMG_W, seed=100).to(device)
    mask = torch.zeros((batch, IMG_H, IMG_W), dtype=torch.bool, device=device)
    with torch.device(device):
        out = dino_ref.dino_forward(sd, cfg, x, mask=mask)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = dino_ref.dino_forward(sd, cfg, x, mask=mask)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"value": round(batch / ms * 1e3, 1), "unit": "images/s", "ms_per_step": round(ms, 2), "batch": batch, "dtype": "f32",
            "what": "oracle/dino_ref.py on CUDA (eager torch fp32, TF32 off) + the reference's own MSDA CUDA kernel (oracle/_ref, sm_100a)"}


def parity_mode_leg(device, dev_imgs, steps=5, exact_steps=2):
    """The parity modes of the SAME engine at the bench shape, device-resident, CUDA-graph replay, timed beside the 16-bit throughput
    mode the headline runs in.  `value`: the split-precision mode (model.split_precision: fp32 activations, every
    Linear / conv a 3-term fp16 split product on the tcgen05 GEMM / implicit-GEMM conv kernels, exact fp32 deformable-attention core) --
    within the north-star 1e-3 of the oracle at this shape (tests/test_gpu_engine.py::test_bench_shape_split_precision_vs_oracle).
    `exact_fp32`: the SIMT fp32 mode (~1e-6, test_bench_shape_fp32_vs_oracle)."""
    def timed(model, n):
        with torch.no_grad():
            for _ in range(2):
                model(dev_imgs)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                model(dev_imgs)
            e1.record()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    B = int(dev_imgs.shape[0])
    model = build_ours(device, torch.float32)
    model.split_precision = True
    ms = timed(model, steps)
    res = {"value": round(B / ms * 1e3, 1), "unit": "images/s", "ms_per_step": round(ms, 2), "batch": B, "dtype": "f32 activations, 2 x f16 split operands",
           "steps": steps,
           "what": "dtlr_b200 engine, compute_dtype = float32 + split_precision: 3 tcgen05 products per fp32 product (hi.hi + hi.lo + lo.hi, "
                   "fp32 accumulation), fp16 tcgen05 self-attention core, exact fp32 MSDA / LayerNorm / GroupNorm / stem; within 1e-3 of the "
                   "oracle on every stage at this shape (measured ~2e-4 on the logits)"}
    try:
        model.split_precision = False
        model.invalidate_engine()
        ms_x = timed(model, exact_steps)
        res["exact_fp32"] = {"value": round(B / ms_x * 1e3, 1), "ms_per_step": round(ms_x, 2), "steps": exact_steps,
                             "what": "the same engine on exact-fp32 SIMT kernels (sgemm_kernel / mha_simt_kernel): ~1e-6 of the oracle; a correctness mode, not tuned"}
    except Exception as e:
        res["exact_fp32"] = {"error": repr(e)[:200]}
    del model
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f16", choices=["f16", "bf16", "f32", "split"],
                    help="f16 (default) / bf16: 16-bit tensor-core operands + activations, fp32 accumulation; f32: exact SIMT parity mode; "
                         "split: fp32 activations, 3-term fp16 split products on the tensor cores (the parity mode at 1e-3, DESIGN.md 3.5b)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-step", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-parity-mode", action="store_true")
    ap.add_argument("--train-impl", default="native", choices=["native", "torch"],
                    help="fine-tune step leg: native TrainEngine (default) or the torch-autograd module path")
    ap.add_argument("--train-ab", action="store_true", help="also time the torch-autograd fine-tune step beside the native one")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    from dtlr_b200 import dist_util
    # The environment's NCCL_DEBUG is left alone (the driver reads NCCL's own rank report).  NCCL logs to file descriptor 1; to keep
    # stdout to the ONE JSON line, fd 1 is pointed at stderr for the run and the line is written to the saved original stdout.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist_util.init("nccl", device)

    from dtlr_b200 import _lib, dino, synth
    dtype = {"f16": torch.float16, "bf16": torch.bfloat16, "f32": torch.float32, "split": torch.float32}[args.dtype]
    model = build_ours(device, dtype)
    model.split_precision = args.dtype == "split"
    B = BATCH_PER_GPU
    host_imgs = synth.synth_images(B, IMG_H, IMG_W, seed=100 + rank).pin_memory()
    dev_imgs = host_imgs.to(device, non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        dist_util.barrier(device)

    def max_over_ranks(ms):
        return dist_util.max_over_ranks(ms, device)

    # ---------------- device-resident throughput (`value`) with the dominant kernel family timed by CUDA events
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    with torch.no_grad():
        for _ in range(args.warmup):
            model(dev_imgs)
    gemm_events, cur = [], {}

    msda_events = []
    ffn_events = []
    buckets = {"gemm": gemm_events, "msda": msda_events, "ffn": ffn_events}

    def timer(name, qty, dev, begin):
        if begin:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
            cur["e0"], cur["fl"] = e0, qty
        else:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            buckets[name].append((cur["e0"], e1, cur["fl"]))

    barrier()
    _lib.LAUNCHES = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.begin()
    e0.record()
    with torch.no_grad():
        for _ in range(args.steps):
            out = model(dev_imgs)
    e1.record()
    barrier()
    sampler.end()
    launches = _lib.LAUNCHES
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = dist_util.whole_job_throughput(B, world, args.steps, ms_total)

    # ---------------- roofline of the dominant kernel family: the same steps once more, launched eagerly (no CUDA graph) with
    # a CUDA-event pair around every dtlr_gemm launch on its stream; durations are per launch, FLOPs are algorithmic 2*M*N*K
    use_graph = model.use_cuda_graph
    model.use_cuda_graph = False
    n_prof = min(args.steps, 3)
    with torch.no_grad():
        model(dev_imgs)
        torch.cuda.synchronize()
        _lib.TIMER = timer
        _lib.GEMM_BYTES = 0
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(n_prof):
            model(dev_imgs)
        p1.record()
        torch.cuda.synchronize()
        _lib.TIMER = None
    model.use_cuda_graph = use_graph
    prof_ms = p0.elapsed_time(p1)
    gemm_ms = sum(a.elapsed_time(b) for a, b, _ in gemm_events)
    gemm_flops = sum(f for _, _, f in gemm_events)
    gemm_bytes = float(getattr(_lib, "GEMM_BYTES", 0))
    n_gemm = len(gemm_events)
    msda_ms = sum(a.elapsed_time(b) for a, b, _ in msda_events)
    ffn_ms = sum(a.elapsed_time(b) for a, b, _ in ffn_events)
    ffn_flops = sum(f for _, _, f in ffn_events)
    msda_bytes = sum(q for _, _, q in msda_events)

    # ---------------- end to end through the public API with HOST buffers (`e2e`)
    from dtlr_b200.misc import nested_tensor_from_tensor_list

    from dtlr_b200.pipeline import HostPipeline
    pipe = HostPipeline(model, device)

    def e2e_run(n):
        last = None
        for ids in pipe.run(host_imgs for _ in range(n)):       # every step re-uploads its batch from pinned host memory
            last = ids
        return last

    with torch.no_grad():
        ids = e2e_run(2)
        barrier()
        sampler.begin()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        ids = e2e_run(args.steps)
        t1.record()
        barrier()
        sampler.end()
    e2e_ms = max_over_ranks(t0.elapsed_time(t1))
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)

    # ---------------- the same end to end with the GPU input stage: the host hands over 8-bit grayscale lines, 1 byte per
    # pixel crosses PCIe, dtlr_preprocess_u8 normalises / pads on the GPU; bucketed LineEvaluator loop, class-id lists back on the host
    import numpy as np
    from dtlr_b200.evaluation import LineEvaluator
    rng = np.random.default_rng(200 + rank)
    u8_lines = [rng.integers(0, 256, (IMG_H, IMG_W), dtype=np.uint8) for _ in range(B)]
    evaluator = LineEvaluator(model, [chr(0x21 + i) for i in range(166)], batch_size=B, width_multiple=32, eps=0.003)
    evaluator.predict(u8_lines * 2)
    barrier()
    sampler.begin()
    u0 = torch.cuda.Event(enable_timing=True); u1 = torch.cuda.Event(enable_timing=True)
    u0.record()
    u8_preds = evaluator.predict(u8_lines * args.steps)
    u1.record()
    barrier()
    sampler.end()
    u8_ms = max_over_ranks(u0.elapsed_time(u1))
    u8_value = world * B * args.steps / (u8_ms / 1e3)
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- BASELINE config 5: the fine-tune step at this N (outside the clock-sampled inference regions)
    train = None
    if not args.no_train_step:
        try:
            train = train_step_leg(device, rank, world, local, impl=args.train_impl)
            if args.train_impl == "native" and args.train_ab:
                try:
                    train["torch_autograd_ab"] = {k: v for k, v in train_step_leg(device, rank, world, local, impl="torch").items()
                                                  if k in ("value", "ms_per_step", "impl")}
                except Exception as e:
                    train["torch_autograd_ab"] = {"error": repr(e)[:200]}
        except Exception as e:      # a secondary leg must never cost the headline line
            train = {"error": repr(e)[:300]}
            try:
                barrier()
            except Exception:
                pass

    if rank != 0:
        dist_util.shutdown()
        return

    pk, pk_kind = peaks()
    peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    ach_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    try:        # secondary view, added after the last GPU run of the round: must never cost the headline line
        hbm_view = gemm_hbm_view(gemm_bytes, gemm_ms, n_prof, pk["hbm_gbs"])
    except Exception as e:
        hbm_view = {"error": repr(e)}
    roofline_gemm = {"kernel": "gemm_ws_tcgen05_kernel + gemm_bf16_tcgen05_kernel (dtlr_gemm: all Linear / 1x1-conv / im2col-conv contractions outside the FFN blocks)" if dtype != torch.float32 else ("gemm_bf16_tcgen05_kernel<.., float> on split operands + dtlr_split_cast (split-precision mode; FLOPs algorithmic, i.e. 1/3 of the tensor-core work)" if args.dtype == "split" else "sgemm_kernel (fp32 parity mode)"),
                     "bound": "tensor", "achieved": round(ach_tf, 2), "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": round(ach_tf / peak_tf, 4), "traffic": None, "peak_kind": pk_kind + " sustained cuBLAS bf16",
                     "launches_timed": n_gemm, "share_of_step": round(gemm_ms / n_prof / ms_step, 3),
                     "algorithmic_flops_per_step": gemm_flops / n_prof,
                     "hbm_view": hbm_view,
                     "note": "K <= 256 for ~150 of these launches (arithmetic intensity <= 128 flop/B): HBM / epilogue bound, not tensor bound -- DESIGN.md 3.2"}
    # the dominant kernel of the step (profiles/r1_launches_step_v6.txt: 18.9 %): the fused FFN block, one launch per encoder /
    # decoder layer; algorithmic FLOPs 4*M*hid*256 per launch; `traffic` = DRAM bytes of one launch from
    # the ncu --set full capture in profiles/r1_ffn_ncu.txt
    ffn_traffic, ffn_traffic_src = profile_traffic(r"ffn_ln_(sk|tcgen05)_kernel", "ffn")
    msda_traffic, msda_traffic_src = profile_traffic(r"msda_fwd", "msda")
    if ffn_events:
        ffn_tf = ffn_flops / (ffn_ms * 1e-3) / 1e12
        roofline = {"kernel": "ffn_ln_sk_kernel<PAIR> (dtlr_ffn_ln_ws: linear1 + ReLU + linear2 + residual + LayerNorm as one stream-K tcgen05 kernel on CTA pairs, cta_group::2; hidden activation in TMEM)",
                    "bound": "tensor", "achieved": round(ffn_tf, 2), "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": round(ffn_tf / peak_tf, 4), "traffic": ffn_traffic, "traffic_from": ffn_traffic_src,
                    "peak_kind": pk_kind + " sustained cuBLAS bf16 (the kernel runs inside a long step)",
                    "launches_timed": len(ffn_events), "share_of_step": round(ffn_ms / n_prof / ms_step, 3),
                    "algorithmic_flops_per_launch": ffn_flops / len(ffn_events),
                    "us_per_launch": round(1e3 * ffn_ms / len(ffn_events), 1),
                    "eager_profiled_ms_per_step": round(prof_ms / n_prof, 3)}
    else:                                   # fp32 parity mode / DTLR_FFN_FUSED=0: the GEMM family is the dominant one
        roofline = dict(roofline_gemm, eager_profiled_ms_per_step=round(prof_ms / n_prof, 3))
    # the north-star kernel (multi-scale deformable attention core, fused prologue): HBM roofline from its algorithmic bytes;
    # `traffic` = dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full, profiles/r1_msda_mma_ncu.txt)
    hbm = pk["hbm_gbs"]
    msda_gbs = msda_bytes / (msda_ms * 1e-3) / 1e9 if msda_ms > 0 else 0.0
    roofline_msda = {"kernel": "msda_fwd_mma_kernel (dtlr_msda_forward_fused)", "bound": "hbm", "achieved": round(msda_gbs, 1), "peak": hbm,
                     "unit": "GB/s", "frac": round(msda_gbs / hbm, 4), "traffic": msda_traffic, "traffic_from": msda_traffic_src, "peak_kind": pk_kind + " copy bandwidth",
                     "launches_timed": len(msda_events), "share_of_step": round(msda_ms / n_prof / ms_step, 3),
                     "algorithmic_bytes_per_launch": round(msda_bytes / max(1, len(msda_events))),
                     "binding_resource": "shared-memory gather bandwidth (4 KB of taps per (query, head) at 128 B/clk/SM), see DESIGN.md 3.1"}
    # the resource that actually binds it: shared-memory wavefronts (one 128-byte wavefront per clock per SM).  20.2 M wavefronts per
    # launch (ncu l1tex__data_pipe_lsu_wavefronts_mem_shared.sum, profiles/r1_msda_mma_ncu.txt) / (SMs x SM clock) = the floor
    try:
        wf, wf_src = profile_metrics(r"msda_fwd", ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",), "msda")
        if msda_events and clocks and clocks.get("sm_mhz") and wf:
            roofline_msda.update(msda_binding_view(1e3 * msda_ms / len(msda_events),
                                                   torch.cuda.get_device_properties(device).multi_processor_count, clocks["sm_mhz"],
                                                   wf["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"], wf_src))
    except Exception as e:      # secondary view, added after the last GPU run of the round
        roofline_msda["binding_error"] = repr(e)
    cpu = None
    if not args.no_cpu_baseline and world == 1:          # rank 0 at N = 1 only; --impl reference gives the N > 1 arm
        ips, sec = cpu_baseline_run(8, 3, 1)
        cpu = {"value": round(ips, 3), "unit": "images/s", "cores": torch.get_num_threads(), "host_cores": os.cpu_count(), "kind": "port",
               "sample": "3 forwards of 8 images (3x40x1024) after 1 warm-up, oracle/dino_ref.py + oracle/msda_ref.c, fp32"}
    gpu_ref = None
    if not args.no_gpu_reference and world == 1:
        try:
            gpu_ref = gpu_reference_leg(device, B)
        except Exception as e:
            gpu_ref = {"error": repr(e)[:300]}
    parity = None
    if not args.no_parity_mode and world == 1 and dtype != torch.float32:
        try:
            parity = parity_mode_leg(device, dev_imgs)
        except Exception as e:      # a secondary leg must never cost the headline line
            parity = {"error": repr(e)[:300]}
    line = {"metric": "text-line images/sec (DINO forward)", "value": round(value, 1), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d (independent shards, no collective)" % world,
                       "weights": "random (dtlr_b200.synth, seed 0)",
                       "precision": {"f16": "fp16 operands + activations (kind::f16 tcgen05 at the bf16 rate: GEMMs, convs, FFN block, decoder self-attention; mma.sync in the MSDA gather), fp32 accumulation; measured vs the fp32 oracle at this shape: see tests/test_gpu_engine.py::test_bench_shape_throughput_mode_vs_oracle and DESIGN.md 2.1",
                                     "bf16": "bf16 operands + activations, fp32 accumulation", "f32": "fp32 SIMT parity mode",
                                     "split": "fp32 activations, every Linear / conv a 3-term fp16 split product on tcgen05 (fp32 accumulation), fp16 self-attention core, exact fp32 MSDA core; within 1e-3 of the oracle at this shape (tests/test_gpu_engine.py::test_bench_shape_split_precision_vs_oracle)"}[args.dtype], "outputs": "all reference dict keys (6 decoder layers + interm)",
                       "l2": "no explicit flush: one step streams >1 GB of activations (126 MB L2)",
                       "launch": "one CUDA-graph replay per step" if model.use_cuda_graph else "eager launches"},
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 1), "unit": "images/s", "h2d_bytes_per_step": host_imgs.numel() * 4 ,
                    "d2h_bytes_per_step": int(ids.numel() * 4), "ms_per_step": round(e2e_ms / args.steps, 3),
                    "api": "dtlr_b200.pipeline.HostPipeline: pinned host images -> DINO.forward -> dino.decode_frames (fused CTC-view argmax) -> pinned host int32 frame ids; H2D / compute / D2H of consecutive steps overlap"},
            "e2e_u8": {"value": round(u8_value, 1), "unit": "images/s", "h2d_bytes_per_step": evaluator.prep.h2d_bytes,
                       "d2h_bytes_per_step": int(ids.numel() * 4), "ms_per_step": round(u8_ms / args.steps, 3),
                       "api": "dtlr_b200.evaluation.LineEvaluator.predict: host u8 grayscale lines -> dtlr_preprocess_u8 (ToTensor + Normalize + pad on the GPU) -> DINO.forward -> fused decode -> host class-id lists"},
            "gpu_launches": launches, "roofline": roofline, "roofline_gemm": roofline_gemm, "roofline_msda": roofline_msda,
            "cpu_baseline": cpu, "gpu_reference": gpu_ref, "parity_mode": parity, "train_step": train}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    dist_util.shutdown()


if __name__ == "__main__":
    main()

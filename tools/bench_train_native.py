"""Fine-tune step (BASELINE config 5, 32 lines per GPU): native TrainEngine (bf16 / fp32) against the torch module path.
Prints one JSON line per variant."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import config, dino, synth, train_engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["bf16", "torch"]
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
x = synth.synth_images(B, 40, 1024, seed=1).cuda()
tg = [{k: v.cuda() for k, v in t.items()} for t in synth.synth_targets(B, 166, seed=1)]


def timed(f, n=5, warm=3):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        r = f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n, r


for w in which:
    model, crit, _ = dino.build_dino(config.latin_ctc_args())
    synth.load_synth_weights(model, 0)
    model = model.cuda().train()
    if w == "torch":
        params = [p for p in model.parameters() if p.requires_grad]
        opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=1e-4)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = crit.loss_CTC(model(x, tg), tg, None, None)["loss_CTC"]
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 0.01)
            opt.step()
            return loss
    else:
        eng = train_engine.TrainEngine(model, lr=1e-5, lr_backbone=1e-10, weight_decay=1e-4, max_norm=0.01,
                                       dtype={"bf16": torch.bfloat16, "f32": torch.float32}[w])

        def step():
            return eng.step(x, tg)
    ms, wall, loss = timed(step)
    mem = torch.cuda.max_memory_allocated() / 2**30
    print(json.dumps({"variant": w, "batch": B, "ms_per_step": round(ms, 2), "wall_ms": round(wall, 2), "images_per_s": round(B / ms * 1e3, 1),
                      "loss": round(float(loss), 4), "peak_mem_GiB": round(mem, 2)}), flush=True)
    if w != "torch" and os.environ.get("DTLR_TRAIN_PROFILE"):
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            step(); step()
            torch.cuda.synchronize()
        rows = sorted(((e.self_device_time_total, e.count, e.key) for e in prof.key_averages()
                       if e.self_device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA), reverse=True)
        tot = sum(r[0] for r in rows)
        print("GPU kernel time per step %.2f ms in %d launches" % (tot / 2e3, sum(r[1] for r in rows) // 2))
        for dt, c, k in rows[:45]:
            print("%9.1f us/step %5d x %5.1f %%  %s" % (dt / 2, c // 2, 100 * dt / tot, k[:140]))
    del model
    torch.cuda.empty_cache()

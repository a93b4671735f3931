"""Where the fine-tune step (BASELINE config 5, 32 lines per GPU) spends its time: torch.profiler kernel table (CUPTI), GPU-busy
time against the wall clock of a step, kernel-launch count.  Usage: python tools/profile_train_step.py [B] > gpurun_out/x.txt"""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import config, dino, synth
from torch.profiler import profile, ProfilerActivity

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
model, crit, _ = dino.build_dino(config.latin_ctc_args())
synth.load_synth_weights(model, 0)
model = model.cuda().train()
params = [p for p in model.parameters() if p.requires_grad]
opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=1e-4)
x = synth.synth_images(B, 40, 1024, seed=1).cuda()
tg = [{k: v.cuda() for k, v in t.items()} for t in synth.synth_targets(B, 166, seed=1)]


def fwd():
    out = model(x, tg)
    return crit.loss_CTC(out, tg, None, None)["loss_CTC"]


def step():
    opt.zero_grad(set_to_none=True)
    loss = fwd()
    loss.backward()
    torch.nn.utils.clip_grad_norm_(params, 0.01)
    opt.step()
    return loss


def timed(f, n=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        r = f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n, r


for _ in range(3):
    step()
ms_step, wall_step, loss = timed(step)
opt.zero_grad(set_to_none=True)
ms_fwd, wall_fwd, l2 = timed(fwd)


def fb():
    opt.zero_grad(set_to_none=True)
    l = fwd(); l.backward(); return l


ms_fb, wall_fb, _ = timed(fb)
print(json.dumps({"batch": B, "step_ms": round(ms_step, 2), "step_wall_ms": round(wall_step, 2), "fwd_ms": round(ms_fwd, 2),
                  "fwd_bwd_ms": round(ms_fb, 2), "loss": float(loss)}))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
ev = prof.key_averages()
tot = sum(getattr(e, "self_device_time_total", 0) for e in ev)
n_k = 0
rows = []
for e in ev:
    dt = getattr(e, "self_device_time_total", 0)
    if dt > 0 and e.device_type == torch.autograd.DeviceType.CUDA:
        rows.append((dt, e.count, e.key)); n_k += e.count
rows.sort(reverse=True)
gpu_busy = sum(r[0] for r in rows)
print("GPU kernel time per step: %.2f ms in %d launches (2 steps profiled: %d)" % (gpu_busy / 2e3, n_k // 2, n_k))
for dt, c, k in rows[:70]:
    print("%9.1f us/step %6d x  %5.1f %%  %s" % (dt / 2, c // 2, 100 * dt / gpu_busy, k[:150]))

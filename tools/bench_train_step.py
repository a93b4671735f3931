"""BASELINE config 5 shape on one GPU: IAM fine-tune step = forward(samples, targets) + loss_CTC + backward (+ AdamW),
batch 32 per GPU, module path (torch autograd + C-ABI deformable attention forward/backward), fp32 or bf16 autocast.
Prints one JSON line; also times the MSDA backward kernel alone at the encoder call shape."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import config, dino, msda, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
model, crit, _ = dino.build_dino(config.latin_ctc_args())
synth.load_synth_weights(model, 0)
model = model.cuda().train()
opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-5)
x = synth.synth_images(B, 40, 1024, seed=1).cuda()
tg = [{k: v.cuda() for k, v in t.items()} for t in synth.synth_targets(B, 166, seed=1)]


def step():
    opt.zero_grad(set_to_none=True)
    out = model(x, tg)
    loss = crit.loss_CTC(out, tg, None, None)["loss_CTC"]
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 0.01)
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 5
for _ in range(n):
    loss = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n

# MSDA backward alone (encoder call shape)
shapes = torch.tensor([(5, 128), (3, 64), (2, 32), (1, 16)], device="cuda")
lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
v = torch.randn(B, 912, 8, 32, device="cuda"); loc = torch.rand(B, 912, 8, 4, 4, 2, device="cuda")
w = torch.softmax(torch.randn(B, 912, 8, 16, device="cuda"), -1).view(B, 912, 8, 4, 4); go = torch.randn(B, 912, 256, device="cuda")
for _ in range(3): msda.ms_deform_attn_backward(v, shapes, lsi, loc, w, go)
torch.cuda.synchronize(); e0.record()
for _ in range(10): msda.ms_deform_attn_backward(v, shapes, lsi, loc, w, go)
e1.record(); torch.cuda.synchronize()
bwd_us = e0.elapsed_time(e1) * 100
for _ in range(3): msda.ms_deform_attn_forward(v, shapes, lsi, loc, w)
torch.cuda.synchronize(); e0.record()
for _ in range(10): msda.ms_deform_attn_forward(v, shapes, lsi, loc, w)
e1.record(); torch.cuda.synchronize()
fwd_us = e0.elapsed_time(e1) * 100
print(json.dumps({"config": "fine-tune step (config/Latin_CTC.py, 40x1024, Q=900+DN, CTC), module path, TF32 matmul", "batch": B,
                  "ms_per_step": round(ms, 2), "images_per_s": round(B / ms * 1e3, 1), "loss": round(float(loss), 4),
                  "msda_bwd_us_fp32": round(bwd_us, 1), "msda_fwd_us_fp32": round(fwd_us, 1)}))

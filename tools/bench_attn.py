import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import ops
B, Q, heads, d = 64, 900, 8, 256
qk = torch.randn(B * Q, 2 * d, device="cuda").bfloat16(); v = torch.randn(B * Q, d, device="cuda").bfloat16()
for _ in range(3): ops.mha_self_attention(qk, d, v, None, B, Q, heads, 32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.mha_self_attention(qk, d, v, None, B, Q, heads, 32)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
fl = 4.0 * B * heads * Q * Q * 32
print(json.dumps({"kernel": "mha_flash_bf16", "us": round(us, 1), "TFLOPs": round(fl / us / 1e6, 1)}))

"""The 3x3 convs of the ResNet-50 stages at the bench batch (B = 64, 40x1024 input) for ncu / timing:
layer1 (10x256, 64 -> 64), layer2 (5x128, 128 -> 128), layer3 (3x64, 256 -> 256), layer4 (2x32, 512 -> 512), plus the stride-2 ones.
usage: python tools/profile_conv.py [time]     (time: CUDA-graph timing table instead of plain launches for ncu)"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import ops

B = 64
dt = torch.float16
cases = [("layer1 3x3 s1", 10, 256, 64, 64, 3, 1), ("layer2 3x3 s1", 5, 128, 128, 128, 3, 1), ("layer3 3x3 s1", 3, 64, 256, 256, 3, 1),
         ("layer4 3x3 s1", 2, 32, 512, 512, 3, 1), ("layer2 3x3 s2", 10, 256, 128, 128, 3, 2), ("layer3 3x3 s2", 5, 128, 256, 256, 3, 2),
         ("layer4 3x3 s2", 3, 64, 512, 512, 3, 2), ("layer2 ds 1x1 s2", 10, 256, 256, 512, 1, 2)]
g = torch.Generator(device="cuda").manual_seed(0)
data = []
for name, H, W, C, Cout, k, s in cases:
    x = torch.randn(B * H * W, C, device="cuda", generator=g).to(dt)
    w = (torch.randn(Cout, k * k * C, device="cuda", generator=g) / (k * k * C) ** 0.5).to(dt)
    b = torch.randn(Cout, device="cuda", generator=g)
    data.append((name, H, W, C, Cout, k, s, x, w, b))
if len(sys.argv) > 1 and sys.argv[1] == "time":
    from gemm_probe_util import timeit
    for name, H, W, C, Cout, k, s, x, w, b in data:
        us = timeit(lambda i: ops.conv2d_nhwc(x, w, b, B, H, W, C, k, k // 2, relu=1, stride=s), iters=10)
        Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
        fl = 2.0 * B * Ho * Wo * Cout * k * k * C
        mb = (x.numel() + B * Ho * Wo * Cout + w.numel()) * 2 / 1e6
        print(json.dumps({"conv": name, "us": round(us, 1), "TFLOPs": round(fl / us / 1e6, 1), "alg_MB": round(mb, 1), "GBs": round(mb / us * 1e3, 1)}), flush=True)
else:
    for _ in range(3):
        for name, H, W, C, Cout, k, s, x, w, b in data:
            ops.conv2d_nhwc(x, w, b, B, H, W, C, k, k // 2, relu=1, stride=s)
    torch.cuda.synchronize()

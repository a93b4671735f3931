"""One eager (no CUDA graph) forward step at the bench configuration, for `ncu` launch lists / full captures.
usage: python tools/profile_step.py [n_forwards]   (all forwards are identical; profile the last one)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from dtlr_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
model = bench.build_ours(torch.device("cuda", 0), torch.float16 if os.environ.get("DTLR_PROFILE_DTYPE", "f16") == "f16" else torch.bfloat16)
model.use_cuda_graph = False
x = synth.synth_images(bench.BATCH_PER_GPU, bench.IMG_H, bench.IMG_W, seed=100).cuda()
with torch.no_grad():
    for _ in range(n):
        model(x)
torch.cuda.synchronize()
print("done")

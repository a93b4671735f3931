"""Micro-benchmark of dtlr_gemm (bf16 tcgen05) at BASELINE config-2 shapes (B=64). CUDA events, > L2 working sets rotate."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dtlr_b200 import ops

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))


def run(M, N, K, iters=20):
    nbuf = 4
    a = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(nbuf)]
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(nbuf)]
    for i in range(4):
        ops.gemm(a[i % nbuf], w, bias, out=out[i % nbuf])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ops.gemm(a[i % nbuf], w, bias, out=out[i % nbuf])
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / iters
    e0.record()
    for i in range(iters):
        torch.addmm(bias.bfloat16(), a[i % nbuf], w.T, out=out[i % nbuf])
    e1.record(); torch.cuda.synchronize()
    us_cublas = e0.elapsed_time(e1) * 1000 / iters
    fl = 2.0 * M * N * K
    by = 2.0 * (M * K + N * K + M * N)
    print(json.dumps({"M": M, "N": N, "K": K, "us": round(us, 1), "TFLOPs": round(fl / us / 1e6, 1),
                      "frac_tensor": round(fl / us / 1e6 / PEAK["bf16_tflops"], 3), "GBs": round(by / us / 1e3, 1),
                      "frac_hbm": round(by / us / 1e3 / PEAK["hbm_gbs"], 3), "cublas_us": round(us_cublas, 1)}))


if __name__ == "__main__":
    if len(sys.argv) > 3:
        run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), iters=3)
        sys.exit(0)
    for (M, N, K) in [(58368, 256, 256), (58368, 384, 256), (58368, 2048, 256), (58368, 256, 2048), (57600, 768, 256),
                      (57600, 256, 512), (40960, 512, 128), (40960, 128, 512)]:
        run(M, N, K)

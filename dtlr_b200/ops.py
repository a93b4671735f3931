"""Python wrappers (torch tensors in, torch tensors out) over the C ABI of include/dtlr_b200.h.
PyTorch only provides device memory and the current stream here; every computation is a libdtlr_b200 kernel."""
import torch

from . import _lib as L


def gemm(a, w, bias=None, residual=None, relu=False, out_dtype=None, out=None):
    """C = act(a @ w.T + bias) (+ residual).  a (M,K) row-major (last-dim stride 1, any row pitch), w (N,K),
    bias fp32 (N) or None, residual (M,N) of the output dtype or None."""
    L.require_cuda(a, w, bias, residual)
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1], (a.shape, w.shape)
    assert a.stride(1) == 1 and w.stride(1) == 1 and a.dtype == w.dtype
    M, K = a.shape
    N = w.shape[0]
    if out_dtype is None:
        out_dtype = a.dtype
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.stride(1) == 1 and out.shape == (M, N) and out.dtype == out_dtype
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
    ldr = 0
    if residual is not None:
        assert residual.dtype == out_dtype and residual.shape == (M, N) and residual.stride(1) == 1
        ldr = residual.stride(0)
    with torch.cuda.device(a.device):
        rc = L.lib().dtlr_gemm(L.ptr(a), a.stride(0), L.ptr(w), w.stride(0), L.ptr(bias) if bias is not None else None,
                               L.ptr(residual) if residual is not None else None, ldr, L.ptr(out), out.stride(0),
                               M, N, K, L.dtype_code(a), L.dtype_code(out), 1 if relu else 0, L.stream_ptr(a.device))
    L.check(rc, "dtlr_gemm")
    return out

"""CPU: the operand layouts of the split-precision mode (engine._split_w against dtlr_split_cast's [hi | hi | lo]) -- a torch emulation
of the kernel's rounding shows that the K' = 3K product is the 3-term split product, for whole-row weights (Linear / 1x1 conv / im2col
GEMM) and for the per-tap layout of the implicit-GEMM 3x3 convs, to ~2^-20 of the fp32 result (fp16 halves) where one fp16 product is
at ~2^-10."""
import pytest
import torch
import torch.nn.functional as F

from dtlr_b200 import engine


def split_cast_emul(x, half):
    hi = x.to(half)
    lo = (x - hi.float()).to(half)
    return torch.cat([hi, hi, lo], -1)


@pytest.mark.parametrize("half,tol", [(torch.float16, 2e-6), (torch.bfloat16, 1e-4)])
def test_whole_row_split_product(half, tol):
    g = torch.Generator().manual_seed(0)
    a = torch.randn(300, 256, generator=g) * 3
    w = torch.randn(384, 256, generator=g) / 16
    ref = a.double() @ w.double().T
    a3 = split_cast_emul(a, half)
    w3 = engine._split_w(w, half)
    assert w3.shape == (384, 768) and w3.dtype == half
    out = a3.double() @ w3.double().T           # exact products of the 16-bit values, as the tensor core forms them
    err = (out - ref).abs().max() / ref.abs().max()
    one = (a.to(half).double() @ w.to(half).double().T - ref).abs().max() / ref.abs().max()
    assert err < tol and err < one / 100, (err.item(), one.item())


@pytest.mark.parametrize("stride", [1, 2])
def test_per_tap_layout_matches_implicit_conv_k_order(stride):
    """the implicit GEMM walks K as (kh, kw, channel block of the 3C-channel pixel): weights [Cout, kh*kw, 3C] with [hi | lo | hi] per tap"""
    half = torch.float16
    g = torch.Generator().manual_seed(1)
    B, C, H, W, Co = 2, 64, 6, 16, 128
    x = torch.randn(B, C, H, W, generator=g)
    wt = torch.randn(Co, C, 3, 3, generator=g) / 24
    ref = F.conv2d(x.double(), wt.double(), padding=1, stride=stride)
    x3 = split_cast_emul(x.permute(0, 2, 3, 1).reshape(-1, C), half).view(B, H, W, 3 * C).permute(0, 3, 1, 2)     # NHWC pixels, 3C channels
    w_rows = wt.permute(0, 2, 3, 1).reshape(Co, 9 * C)                                                             # engine._fold_conv_bn order
    w3 = engine._split_w(w_rows, half, taps=9).view(Co, 3, 3, 3 * C).permute(0, 3, 1, 2)
    out = F.conv2d(x3.double(), w3.double(), padding=1, stride=stride)
    err = (out - ref).abs().max() / ref.abs().max()
    assert err < 2e-6, err.item()


def test_split_dtype_is_a_distinct_pack_key():
    a, b = engine.SplitDtype(torch.float16), engine.SplitDtype(torch.float16)
    assert a == b and hash(a) == hash(b) and a != engine.SplitDtype(torch.bfloat16)
    assert a != torch.float32 and not (a in (torch.bfloat16, torch.float16))

"""GPU: the kernels of the split-precision mode -- dtlr_split_cast (bit-exact against its torch restatement), the mixed fp32 x split-weight
dtlr_gemm, and the 3x3 convolution over 3C-channel split pixels (implicit GEMM with an fp32 result, and its im2col form) -- against
fp64 torch references of the SAME fp32 operands.  Floating point: a split product carries 2 x 11 significand bits (fp16 halves), so the
bound asserted is 5e-5 relative-to-max (and at least 3x below what ONE 16-bit product of the same operands gives; bf16 halves: 5e-4)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
HALVES = [(torch.float16, 5e-5), (torch.bfloat16, 5e-4)]


def _split_emul(x, half):
    hi = x.to(half)
    return torch.cat([hi, hi, (x - hi.float()).to(half)], -1)


@pytest.mark.parametrize("half", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,K", [(1, 8), (257, 72), (1800, 256), (130, 2048)])
def test_split_cast_bit_exact(half, M, K):
    from dtlr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + K)
    x = torch.randn(M, K, device="cuda", generator=g) * torch.logspace(-6, 3, K, device="cuda")       # sub-normal lo halves included
    assert torch.equal(ops.split_cast(x, half), _split_emul(x, half))
    wide = torch.randn(M, K + 24, device="cuda", generator=g)
    view = wide[:, 8:8 + K]                                                                            # pitched rows, 16-byte aligned start
    assert torch.equal(ops.split_cast(view, half), _split_emul(view, half))


@pytest.mark.parametrize("half,tol", HALVES)
@pytest.mark.parametrize("M,N,K", [(1824, 256, 256), (1800, 2048, 256), (1824, 256, 2048), (900, 384, 256), (912, 166, 256), (900, 4, 256),
                                   (130, 64, 64), (257, 200, 72), (58368, 256, 64), (40000, 512, 256)])
def test_split_gemm_vs_fp64(half, tol, M, N, K):
    from dtlr_b200 import engine, ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    w3 = engine._split_w(w, half)
    ref = a.double() @ w.double().T + bias.double()
    out = ops.gemm(a, w3, bias)
    assert out.dtype == torch.float32
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    one = ((a.to(half).double() @ w.to(half).double().T + bias.double()) - ref).abs().max().item() / ref.abs().max().item()
    print("split gemm %s M=%d N=%d K=%d: rel-to-max %.2e (one 16-bit product: %.2e)" % (half, M, N, K, err, one))
    assert err < tol and err < one / 3, (err, one)
    res = torch.randn(M, N, device="cuda", generator=g)
    out2 = ops.gemm(a, w3, bias, residual=res, relu=2)
    ref2 = torch.relu(ref + res.double())
    assert (out2.double() - ref2).abs().max().item() / ref2.abs().max().item() < tol
    if N % 4:                                                   # pitched fp32 output (the 166-class heads of the engine)
        buf = torch.empty((M, (N + 3) // 4 * 4), dtype=torch.float32, device="cuda")
        out3 = ops.gemm(a, w3, bias, out_dtype=torch.float32, out=buf[:, :N])
        assert (out3.double() - ref).abs().max().item() / ref.abs().max().item() < tol


@pytest.mark.parametrize("half,tol", HALVES)
@pytest.mark.parametrize("B,C,H,W,Co,stride", [(2, 64, 10, 256, 64, 1), (2, 128, 10, 256, 128, 2), (3, 256, 5, 128, 256, 2),
                                               (2, 512, 3, 64, 512, 2), (2, 512, 2, 32, 512, 1), (1, 64, 6, 40, 64, 1)])
def test_split_conv3x3_vs_fp64(half, tol, B, C, H, W, Co, stride):
    """ResNet-50 3x3 convs (torchvision Bottleneck v1.5, reference models/dino/backbone.py:118-120) in the split mode: NHWC fp32 -> split
    pixels (3C channels) -> implicit GEMM (per-tap [hi | lo | hi] weights) with ReLU and an fp32 result; widths the implicit kernel cannot
    tile take the 16-bit im2col + GEMM form of the same product."""
    from dtlr_b200 import engine, ops
    g = torch.Generator(device="cuda").manual_seed(C + H + W + stride)
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    wt = torch.randn(Co, C, 3, 3, device="cuda", generator=g) / (9 * C) ** 0.5
    bias = torch.randn(Co, device="cuda", generator=g)
    ref = torch.relu(F.conv2d(x.double(), wt.double(), bias.double(), padding=1, stride=stride)).permute(0, 2, 3, 1).reshape(-1, Co)
    a = x.permute(0, 2, 3, 1).reshape(B * H * W, C).contiguous()
    w3 = engine._split_w(wt.permute(0, 2, 3, 1).reshape(Co, 9 * C), half, taps=9)
    a3 = ops.split_cast(a, half)
    outs = {}
    if ops.conv2d_nhwc_supported(a3, H, W, 3 * C, 3, stride):
        outs["implicit"] = ops.conv2d_nhwc(a3, w3, bias, B, H, W, 3 * C, 3, 1, relu=1, stride=stride, out_dtype=torch.float32)[0]
    col = ops.im2col(a3, B, H, W, 3 * C, 3, 3, stride, 1, half)[0]
    outs["im2col"] = ops.gemm(col, w3, bias, relu=1, out_dtype=torch.float32)
    assert "implicit" in outs or 128 % min(128, (W - 1) // stride + 1) != 0        # (output rows of 40 pixels do not tile into 128-pixel segments)
    for name, out in outs.items():
        assert out.dtype == torch.float32 and out.shape == ref.shape
        err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
        print("split conv %s %s C=%d %dx%d s%d: rel-to-max %.2e" % (name, half, C, H, W, stride, err))
        assert err < tol, (name, err)


@pytest.mark.parametrize("half", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K", [(1824, 256, 256), (1800, 2048, 256), (1824, 256, 2048), (900, 384, 256), (130, 64, 64), (257, 200, 72),
                                   (58368, 64, 64)])
def test_split_output_epilogue_bit_exact(half, M, N, K):
    """out_dtype = ops.SPLIT (DTLR_SPLIT16): the GEMM epilogue writes [hi | hi | lo] of its fp32 result -- bit-identical to dtlr_split_cast
    of the fp32 output of the same product (bias, ReLU before / after an fp32 residual), from fp32 and from already-split A operands"""
    from dtlr_b200 import engine, ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K + 1)
    a = torch.randn(M, K, device="cuda", generator=g)
    w3 = engine._split_w(torch.randn(N, K, device="cuda", generator=g) / K ** 0.5, half)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    a3 = ops.split_cast(a, half)
    for kw in (dict(relu=1), dict(relu=2, residual=res), dict()):
        for s3 in (False, True):      # plain walk over 3K columns / hi and lo tiles loaded once (in_dtype DTLR_SPLIT16): same epilogue
            want = ops.split_cast(ops.gemm(a3, w3, bias, out_dtype=torch.float32, split3=s3, **kw), half)
            got = ops.gemm(a3, w3, bias, out_dtype=ops.SPLIT, split3=s3, **kw)
            assert got.shape == (M, 3 * N) and got.dtype == half
            assert torch.equal(got, want), (kw, s3)
        assert torch.equal(ops.gemm(a, w3, bias, out_dtype=ops.SPLIT, **kw), want), kw          # fp32 A: split pass + split3 product
    f_plain = ops.gemm(a3, w3, bias, out_dtype=torch.float32)
    f_s3 = ops.gemm(a3, w3, bias, out_dtype=torch.float32, split3=True)
    d = (f_plain - f_s3).abs().max().item() / f_plain.abs().max().item()
    assert d < 3e-5, d                # the two schedules differ only in the order of the fp32 accumulation (measured 5.8e-6 at K = 2048)


@pytest.mark.parametrize("B,C,H,W,Co,stride", [(2, 64, 10, 256, 64, 1), (2, 128, 10, 256, 128, 2), (2, 512, 2, 32, 512, 1)])
def test_split_output_conv_bit_exact(B, C, H, W, Co, stride):
    from dtlr_b200 import engine, ops
    half = torch.float16
    g = torch.Generator(device="cuda").manual_seed(C + W)
    a3 = ops.split_cast(torch.randn(B * H * W, C, device="cuda", generator=g), half)
    w3 = engine._split_w(torch.randn(Co, 9 * C, device="cuda", generator=g) / (9 * C) ** 0.5, half, taps=9)
    bias = torch.randn(Co, device="cuda", generator=g)
    f32, Ho, Wo = ops.conv2d_nhwc(a3, w3, bias, B, H, W, 3 * C, 3, 1, relu=1, stride=stride, out_dtype=torch.float32)
    got = ops.conv2d_nhwc(a3, w3, bias, B, H, W, 3 * C, 3, 1, relu=1, stride=stride, out_dtype=ops.SPLIT)[0]
    assert got.shape == (B * Ho * Wo, 3 * Co) and torch.equal(got, ops.split_cast(f32, half))

"""GPU: convolution kernels against torch conv2d (floating-point kernels -> plain torch fp32 reference):
implicit-GEMM stride-1 3x3 conv on NHWC bf16 (dtlr_conv2d_nhwc, TMA taps with zero fill), the im2col+GEMM path for strided
convs, the direct 7x7 stem, the max-pool and GroupNorm."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False          # the torch reference convs must be true fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


@pytest.mark.parametrize("B,H,W,C,Cout", [(2, 10, 256, 64, 64), (3, 5, 128, 128, 128), (3, 3, 64, 256, 256), (5, 2, 32, 512, 512), (2, 1, 16, 64, 128)])
def test_implicit_gemm_conv3x3(B, H, W, C, Cout):
    from dtlr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B * H + W)
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, C, 3, 3, device="cuda", generator=g) / (9 * C) ** 0.5
    bias = torch.randn(Cout, device="cuda", generator=g)
    xb = x.permute(0, 2, 3, 1).reshape(B * H * W, C).bfloat16().contiguous()
    wb = w.permute(0, 2, 3, 1).reshape(Cout, 9 * C).bfloat16().contiguous()
    assert ops.conv2d_nhwc_supported(xb, H, W, C, 3, 1)
    out, Ho, Wo = ops.conv2d_nhwc(xb, wb, bias, B, H, W, C, 3, 1, relu=1)
    assert (Ho, Wo) == (H, W)
    ref = F.relu(F.conv2d(xb.float().view(B, H, W, C).permute(0, 3, 1, 2), wb.float().view(Cout, 3, 3, C).permute(0, 3, 1, 2), bias, padding=1))
    ref = ref.permute(0, 2, 3, 1).reshape(B * H * W, Cout)
    assert (out.float() - ref).abs().max().item() / ref.abs().max().item() < 2e-2
    res = torch.randn(B * H * W, Cout, device="cuda", generator=g).bfloat16()
    out2 = ops.conv2d_nhwc(xb, wb, bias, B, H, W, C, 3, 1, relu=2, residual=res)[0]
    ref2 = F.relu(F.conv2d(xb.float().view(B, H, W, C).permute(0, 3, 1, 2), wb.float().view(Cout, 3, 3, C).permute(0, 3, 1, 2), bias, padding=1)
                  .permute(0, 2, 3, 1).reshape(B * H * W, Cout) + res.float())
    assert (out2.float() - ref2).abs().max().item() / ref2.abs().max().item() < 2e-2


@pytest.mark.parametrize("B,H,W,C,Cout,k", [(2, 10, 256, 128, 128, 3), (3, 5, 128, 256, 256, 3), (5, 3, 64, 512, 512, 3), (64, 10, 256, 128, 128, 3),
                                            (2, 10, 256, 256, 512, 1), (3, 5, 128, 512, 1024, 1), (4, 3, 64, 1024, 2048, 1), (2, 4, 64, 64, 64, 3)])
def test_strided_implicit_gemm_conv(B, H, W, C, Cout, k):
    """stride-2 'same'-padded convs of the first Bottleneck of layer2-4 (3x3 conv2 and the 1x1 downsample) as implicit GEMMs: the A
    tensor map traverses W and H with element stride 2 (dtlr_conv2d_nhwc_strided) -- against torch conv2d and, exactly, against the
    im2col + GEMM path it replaces (odd heights: 5 -> 3, 3 -> 2 with the bottom padding row)"""
    from dtlr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B * H + W + k)
    pad = k // 2
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, C, k, k, device="cuda", generator=g) / (k * k * C) ** 0.5
    bias = torch.randn(Cout, device="cuda", generator=g)
    xb = x.permute(0, 2, 3, 1).reshape(B * H * W, C).bfloat16().contiguous()
    wb = w.permute(0, 2, 3, 1).reshape(Cout, k * k * C).bfloat16().contiguous()
    assert ops.conv2d_nhwc_supported(xb, H, W, C, k, 2)
    out, Ho, Wo = ops.conv2d_nhwc(xb, wb, bias, B, H, W, C, k, pad, relu=1, stride=2)
    ref = F.relu(F.conv2d(xb.float().view(B, H, W, C).permute(0, 3, 1, 2), wb.float().view(Cout, k, k, C).permute(0, 3, 1, 2), bias, stride=2, padding=pad))
    assert (Ho, Wo) == tuple(ref.shape[-2:])
    ref = ref.permute(0, 2, 3, 1).reshape(B * Ho * Wo, Cout)
    assert (out.float() - ref).abs().max().item() / ref.abs().max().item() < 2e-2
    col, Ho2, Wo2 = ops.im2col(xb, B, H, W, C, k, k, 2, pad, torch.bfloat16)
    old = ops.gemm(col, wb, bias, relu=1)
    assert (Ho2, Wo2) == (Ho, Wo) and (out.float() - old.float()).abs().max().item() <= 2 ** -7 * ref.abs().max().item()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_strided_conv_im2col_gemm(dtype):
    from dtlr_b200 import ops
    B, H, W, C, Cout = 2, 10, 256, 128, 128
    x = torch.randn(B, C, H, W, device="cuda")
    w = torch.randn(Cout, C, 3, 3, device="cuda") / (9 * C) ** 0.5
    xb = x.permute(0, 2, 3, 1).reshape(B * H * W, C).to(dtype).contiguous()
    wb = w.permute(0, 2, 3, 1).reshape(Cout, 9 * C).to(dtype).contiguous()
    col, Ho, Wo = ops.im2col(xb, B, H, W, C, 3, 3, 2, 1, dtype)
    out = ops.gemm(col, wb, torch.zeros(Cout, device="cuda"))
    ref = F.conv2d(xb.float().view(B, H, W, C).permute(0, 3, 1, 2), wb.float().view(Cout, 3, 3, C).permute(0, 3, 1, 2), stride=2, padding=1)
    assert (Ho, Wo) == tuple(ref.shape[-2:])
    ref = ref.permute(0, 2, 3, 1).reshape(B * Ho * Wo, Cout)
    assert (out.float() - ref).abs().max().item() / ref.abs().max().item() < (1e-5 if dtype == torch.float32 else 2e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("H,W", [(40, 1024), (40, 704), (37, 133)])
def test_stem_conv_and_maxpool(dtype, H, W):
    from dtlr_b200 import ops
    B = 2
    x = torch.randn(B, 3, H, W, device="cuda")
    w = torch.randn(64, 3, 7, 7, device="cuda") / 147 ** 0.5
    bias = torch.randn(64, device="cuda")
    y, Ho, Wo = ops.stem_conv(x, w.permute(2, 3, 1, 0).contiguous(), bias, B, H, W, dtype)
    ref = F.relu(F.conv2d(x, w, bias, stride=2, padding=3))
    assert (Ho, Wo) == tuple(ref.shape[-2:])
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert (y.float().view(B, Ho, Wo, 64).permute(0, 3, 1, 2) - ref).abs().max().item() / ref.abs().max().item() < tol
    p, Hp, Wp = ops.maxpool3x3s2(y, B, Ho, Wo, 64)
    refp = F.max_pool2d(y.float().view(B, Ho, Wo, 64).permute(0, 3, 1, 2), 3, 2, 1)
    assert (Hp, Wp) == tuple(refp.shape[-2:])
    assert torch.equal(p.float().view(B, Hp, Wp, 64).permute(0, 3, 1, 2), refp)


def test_groupnorm_into_token_buffer():
    from dtlr_b200 import ops
    B, HW, C, S, off = 3, 40, 256, 100, 17
    x = torch.randn(B * HW, C, device="cuda") * 3 + 1
    gw, gb = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    buf = torch.zeros(B * S, C, device="cuda")
    ops.groupnorm_into(x, gw, gb, buf, B, HW, C, 32, off, S)
    ref = F.group_norm(x.view(B, HW, C).permute(0, 2, 1), 32, gw, gb, 1e-5).permute(0, 2, 1)
    got = buf.view(B, S, C)[:, off:off + HW]
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4)
    assert buf.view(B, S, C)[:, :off].abs().sum() == 0 and buf.view(B, S, C)[:, off + HW:].abs().sum() == 0


@pytest.mark.parametrize("B,H,W", [(2, 40, 1024), (3, 37, 133), (64, 40, 1024)])
def test_stem_as_staged_im2col_plus_tensor_core_gemm(B, H, W):
    """bf16 throughput mode stem: stem_im2col_kernel (fp32 NCHW -> bf16 (kh,kw,c) patches, K padded 147 -> 152) + dtlr_gemm
    vs torch conv2d on the same bf16-rounded operands, including ragged widths (133) and odd heights."""
    from dtlr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B + H + W)
    x = torch.randn(B, 3, H, W, device="cuda", generator=g)
    w = torch.randn(64, 3, 7, 7, device="cuda", generator=g) / 147 ** 0.5
    bias = torch.randn(64, device="cuda", generator=g)
    wk = F.pad(w.permute(0, 2, 3, 1).reshape(64, 147), (0, 5)).bfloat16().contiguous()
    col, Ho, Wo = ops.im2col(x, B, H, W, 3, 7, 7, 2, 3, torch.bfloat16, nchw_input=True, ldo=152)
    # the patch matrix itself is exact (bf16 rounding of the input only): compare with unfold
    ref_col = F.unfold(x.bfloat16().float(), 7, padding=3, stride=2).view(B, 3, 49, Ho * Wo).permute(0, 3, 2, 1).reshape(B * Ho * Wo, 147)
    assert torch.equal(col[:, :147].float(), ref_col) and col[:, 147:].abs().sum() == 0
    y = ops.gemm(col, wk, bias, relu=1)
    ref = F.relu(F.conv2d(x.bfloat16().float(), w.bfloat16().float(), bias, stride=2, padding=3))
    assert (Ho, Wo) == tuple(ref.shape[-2:])
    err = (y.float().view(B, Ho, Wo, 64).permute(0, 3, 1, 2) - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-2, err

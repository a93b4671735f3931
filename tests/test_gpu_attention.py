"""GPU: decoder self-attention kernels (dtlr_mha_self_attention) against a plain torch fp32 reference of the same op
(floating-point kernel): exact-fp32 SIMT path (with and without the boolean attn_mask of training mode) and the bf16
tensor-core flash path."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def ref_attention(qk, v, B, Q, heads, mask=None):
    d = v.shape[1]
    dh = d // heads
    q = qk[:, :d].float().view(B, Q, heads, dh).transpose(1, 2)
    k = qk[:, d:].float().view(B, Q, heads, dh).transpose(1, 2)
    vv = v.float().view(B, Q, heads, dh).transpose(1, 2)
    s = (q / math.sqrt(dh)) @ k.transpose(-1, -2)
    if mask is not None:
        s = s.masked_fill(mask[None, None], float("-inf"))
    return (torch.softmax(s, -1) @ vv).transpose(1, 2).reshape(B * Q, d)


@pytest.mark.parametrize("Q", [900, 986, 100, 17, 1024, 912, 130, 65, 8, 513])
@pytest.mark.parametrize("two_pass", [False, True])
def test_self_attention_tcgen05_kernel(Q, two_pass, monkeypatch):
    """the tcgen05 / TMEM / TMA attention kernels (the default without a mask; DTLR_ATTN) against torch fp32 on the same bf16
    operands: the single-pass kernel (4 tiles in flight, P in tensor memory, online softmax with TMEM rescale; Q covers query tiles
    with dead warps, 16-key / 1-key / full last chunks, one to eight tiles) and the two-pass kernel (dtlr_debug_flags(256))"""
    from dtlr_b200 import ops, _lib
    monkeypatch.setattr(ops, "ATTN_IMPL", "tc")
    _lib.lib().dtlr_debug_flags(256 if two_pass else 0)
    B, heads, d = 3, 8, 256
    g = torch.Generator(device="cuda").manual_seed(Q + 1)
    qk = (torch.randn(B * Q, 2 * d, device="cuda", generator=g) * 1.5).bfloat16()
    v = torch.randn(B * Q, d, device="cuda", generator=g).bfloat16()
    before = ops.L.LAUNCHES
    out = ops.mha_self_attention(qk, d, v, None, B, Q, heads, d // heads)
    assert ops.L.LAUNCHES - before == 2          # transpose pre-pass + tcgen05 kernel, not the fallback
    _lib.lib().dtlr_debug_flags(0)
    ref = ref_attention(qk, v, B, Q, heads)
    assert (out.float() - ref).abs().max().item() / ref.abs().max().item() < 2e-2


@pytest.mark.parametrize("Q", [900, 986, 100, 17])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_self_attention(Q, dtype):
    from dtlr_b200 import ops
    B, heads, d = 3, 8, 256
    g = torch.Generator(device="cuda").manual_seed(Q)
    qk = (torch.randn(B * Q, 2 * d, device="cuda", generator=g) * 1.5).to(dtype)
    v = torch.randn(B * Q, d, device="cuda", generator=g).to(dtype)
    out = ops.mha_self_attention(qk, d, v, None, B, Q, heads, d // heads)
    ref = ref_attention(qk, v, B, Q, heads)
    err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
    assert err < (1e-5 if dtype == torch.float32 else 2e-2), err


def test_self_attention_with_training_mask():
    from dtlr_b200 import ops
    B, Q, heads, d, pad = 2, 386, 8, 256, 86
    g = torch.Generator(device="cuda").manual_seed(1)
    qk = torch.randn(B * Q, 2 * d, device="cuda", generator=g)
    v = torch.randn(B * Q, d, device="cuda", generator=g)
    mask = torch.zeros(Q, Q, dtype=torch.bool, device="cuda")
    mask[pad:, :pad] = True
    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 2e-2)):
        out = ops.mha_self_attention(qk.to(dtype), d, v.to(dtype), mask.view(torch.uint8), B, Q, heads, d // heads)
        ref = ref_attention(qk.to(dtype), v.to(dtype), B, Q, heads, mask)
        assert (out.float() - ref).abs().max().item() / ref.abs().max().item() < tol

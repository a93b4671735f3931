"""TEST INFRASTRUCTURE -- plain-torch (CPU) stand-ins for the leaf functions of dtlr_b200/ops.py and dtlr_b200/msda.py that the fused
inference engine (dtlr_b200/engine.py) launches, so that the engine's HOST LOGIC -- weight packing (BN folding, split-precision layouts),
level geometry, which kernel consumes which buffer in which layout, the fp32 and split-precision orchestration -- runs in the CPU suite
against the reference-generated fixtures.  It is NOT a fallback: nothing under dtlr_b200/ imports it, and `install()` only works through
pytest's monkeypatch.  Each function states the contract of the C-ABI entry point it stands in for (include/dtlr_b200.h)."""
import contextlib

import torch
import torch.nn.functional as F

from dtlr_b200 import _lib as L
from dtlr_b200 import msda as msda_mod
from dtlr_b200 import ops
from dtlr_b200.misc import inverse_sigmoid
from oracle import dino_ref

import train_ops_double as tdb          # msda_prep / msda_core / enc_ref_points / sine_embed restatements shared with the train-engine double

HALF = ops.HALF
CALLS = {}


def _count(name):
    CALLS[name] = CALLS.get(name, 0) + 1


def split_cast(x, dtype=torch.float16):
    """dtlr_split_cast: fp32 (M,K) -> 16-bit (M,3K) = [hi | hi | lo]"""
    _count("split_cast")
    assert x.dtype == torch.float32 and x.dim() == 2 and x.shape[1] % 8 == 0
    hi = x.to(dtype)
    return torch.cat([hi, hi, (x - hi.float()).to(dtype)], 1)


def gemm(a, w, bias=None, residual=None, relu=False, out_dtype=None, out=None, split3=False):
    """dtlr_gemm: C = act(A W^T + bias) (+ residual); relu 1 before / 2 after the residual add; fp32 A against a 16-bit W of 3K columns =
    split product; out_dtype ops.SPLIT = the fp32 result as [hi | hi | lo]; split3 = both operands are whole-row split matrices"""
    _count("gemm")
    if a.dtype == torch.float32 and w.dtype in HALF:
        assert w.shape[1] == 3 * a.shape[1], (a.shape, w.shape)
        out_dtype = torch.float32 if out_dtype is None else out_dtype
        a = split_cast(a, w.dtype)
        split3 = True
    assert a.dtype == w.dtype and a.dim() == 2 and a.shape[1] == w.shape[1], (a.dtype, w.dtype, a.shape, w.shape)
    if split3:       # the kernel loads A_hi at column j and A_lo at 2K + j, W_hi at j and W_lo at K + j: only valid for whole-row layouts
        K = a.shape[1] // 3
        assert a.dtype in HALF and a.shape[1] % 3 == 0 and torch.equal(a[:, :K], a[:, K:2 * K]) and torch.equal(w[:, :K], w[:, 2 * K:])
    relu = int(relu)
    c = (a.double() @ w.double().t()).float()
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == w.shape[0]
        c = c + bias
    if relu == 1:
        c = F.relu(c)
    if residual is not None:
        assert residual.shape == c.shape and residual.dtype == (torch.float32 if out_dtype is ops.SPLIT else (out_dtype or a.dtype))
        c = c + residual.float()
    if relu == 2:
        c = F.relu(c)
    if out_dtype is ops.SPLIT:
        assert out is None and a.dtype in HALF
        return split_cast(c, a.dtype)
    c = c.to(out_dtype or a.dtype)
    if out is not None:
        assert out.shape == c.shape and out.dtype == c.dtype
        out.copy_(c)
        return out
    return c


def im2col(x, B, H, W, C, KH, KW, stride, pad, out_dtype, nchw_input=False, ldo=None):
    """dtlr_im2col: NHWC rows [B*H*W, C] (or the NCHW fp32 image) -> [B*Ho*Wo, KH*KW*C] in (kh, kw, c) order, zero padded to ldo columns"""
    _count("im2col")
    img = x.view(B, C, H, W) if nchw_input else x.view(B, H, W, C).permute(0, 3, 1, 2)
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    u = F.unfold(img.float(), (KH, KW), padding=pad, stride=stride)                         # (B, C*KH*KW, L), channel-major
    u = u.view(B, C, KH * KW, Ho * Wo).permute(0, 3, 2, 1).reshape(B * Ho * Wo, KH * KW * C)
    if ldo and ldo > u.shape[1]:
        u = F.pad(u, (0, ldo - u.shape[1]))
    return u.to(out_dtype).contiguous(), Ho, Wo


def conv2d_nhwc(x, w, bias, B, H, W, C, k, pad, relu=0, residual=None, stride=1, out_dtype=None):
    """dtlr_conv2d_nhwc_strided: the implicit GEMM = im2col in (kh, kw, channel) order x w^T, never materialised on the GPU"""
    _count("conv2d_nhwc")
    assert x.dtype in HALF and ops.conv2d_nhwc_supported(x, H, W, C, k, stride)
    col, Ho, Wo = im2col(x, B, H, W, C, k, k, stride, pad, x.dtype)
    return gemm(col, w, bias, residual=residual, relu=relu, out_dtype=out_dtype), Ho, Wo


def stem_conv(x, w_khkwcico, bias, B, H, W, out_dtype):
    """dtlr_stem_conv: conv1 7x7 / 2 / 3 + folded FrozenBN + ReLU, NCHW fp32 image -> NHWC rows"""
    _count("stem_conv")
    y = F.relu(F.conv2d(x.view(B, 3, H, W), w_khkwcico.permute(3, 2, 0, 1), bias, stride=2, padding=3))
    return y.permute(0, 2, 3, 1).reshape(-1, 64).to(out_dtype).contiguous(), y.shape[2], y.shape[3]


def maxpool3x3s2(x, B, H, W, C):
    _count("maxpool")
    y = F.max_pool2d(x.view(B, H, W, C).permute(0, 3, 1, 2).float(), 3, 2, 1)
    return y.permute(0, 2, 3, 1).reshape(-1, C).to(x.dtype).contiguous(), y.shape[2], y.shape[3]


def groupnorm_into(x_f32, gamma, beta, out, B, HW, C, G, row_offset, rows_per_batch, eps=1e-5):
    """dtlr_groupnorm: nn.GroupNorm(G, C) of one level, written into rows [row_offset, row_offset + HW) of every image of the token buffer"""
    _count("groupnorm")
    y = F.group_norm(x_f32.view(B, HW, C).permute(0, 2, 1), G, gamma, beta, eps).permute(0, 2, 1)
    out.view(B, rows_per_batch, C)[:, row_offset:row_offset + HW] = y.to(out.dtype)


def pos_sine_into(mask_u8, level_embed, out, B, H, W, npf, temp_h, temp_w, row_offset, rows_per_batch):
    """dtlr_pos_sine: PositionEmbeddingSineHW + level_embed, same destination rule as dtlr_groupnorm"""
    _count("pos_sine")
    pe = dino_ref.pos_sine_hw(mask_u8.view(B, H, W).bool(), temp_h, temp_w, npf)           # (B, C, H, W)
    out.view(B, rows_per_batch, 2 * npf)[:, row_offset:row_offset + H * W] = (pe.flatten(2).permute(0, 2, 1) + level_embed).to(out.dtype)


def add_layernorm(x, res, gamma, beta, add2=None, eps=1e-5, out=None):
    _count("add_layernorm")
    z = x.float() if res is None else x.float() + res.float()
    y = F.layer_norm(z, (x.shape[1],), gamma, beta, eps).to(x.dtype)
    if out is not None:
        out.copy_(y)
        y = out
    return (y, (y.float() + add2.float()).to(x.dtype)) if add2 is not None else y


def add(a, b):
    _count("add")
    return a + b


def zero_masked_rows_(x, rowmask_u8):
    _count("zero_masked_rows")
    x[rowmask_u8.bool()] = 0
    return x


def _geo(shapes_host, n_levels, B, S=None):
    v = [int(t) for t in shapes_host]
    return {"B": B, "S": S, "nlev": n_levels, "level_hw": [(v[2 * l], v[2 * l + 1]) for l in range(n_levels)]}


def enc_ref_points(valid_ratios, shapes_host, n_levels, B, S):
    _count("enc_ref_points")
    return tdb.enc_ref_points(valid_ratios, _geo(shapes_host, n_levels, B, S))


def encoder_proposals(memory, pad_u8, valid_hw_i32, shapes_host, n_levels, B, S, C, default_hw):
    """dtlr_encoder_proposals = gen_encoder_output_proposals (reference models/dino/utils.py:15-64)"""
    _count("encoder_proposals")
    shapes = torch.tensor(_geo(shapes_host, n_levels, B)["level_hw"])
    mem, prop = dino_ref.gen_encoder_output_proposals(memory.view(B, S, C).float(), pad_u8.view(B, S).bool(), shapes, default_hw)
    return mem.reshape(B * S, C).to(memory.dtype), prop.reshape(B * S, 4)


def rowmax(x_f32, N):
    _count("rowmax")
    return x_f32[:, :N].max(1)[0]


def sine_embed(ref, valid_ratios, B, Q, n_levels, out_dtype):
    _count("sine_embed")
    return tdb.sine_embed(ref, valid_ratios, B, Q, n_levels, out_dtype)


def box_refine(delta_f32, ref):
    _count("box_refine")
    return (delta_f32[:, :4] + inverse_sigmoid(ref)).sigmoid()


def cast(x, dtype):
    _count("cast")
    return x.to(dtype)


def mha_self_attention(qk, k_off, v, attn_mask_u8, B, Q, heads, head_dim):
    """dtlr_mha_self_attention / dtlr_mha_tcgen05: softmax(q k^T / sqrt(d)) v per (image, head); q = qk[:, :k_off], k = qk[:, k_off:]"""
    _count("mha")
    C = heads * head_dim
    q, k = qk[:, :C].float(), qk[:, k_off:k_off + C].float()

    def sp(t):
        return t.view(B, Q, heads, head_dim).transpose(1, 2)
    s = (sp(q) / head_dim ** 0.5) @ sp(k).transpose(-1, -2)
    if attn_mask_u8 is not None:
        s = s.masked_fill(attn_mask_u8.bool()[None, None], float("-inf"))
    o = (F.softmax(s, -1) @ sp(v.float())).transpose(1, 2).reshape(B * Q, C)
    return o.to(v.dtype)


def topk_select(scores, K):
    _count("topk_select")
    return torch.topk(scores, K, dim=1)[1]


def select_gather(idx, coord, prop, mem):
    _count("select_gather")
    B, K = idx.shape

    def g(t):
        return torch.gather(t, 1, idx[..., None].expand(B, K, t.shape[-1]))
    return g((coord + prop).sigmoid()), g(prop.sigmoid()), g(mem)


def msda_forward_fused(value, shapes_host, lsi_host, n_levels, proj, ref, valid_ratios, Lq, P, out=None):
    """dtlr_msda_forward_fused: softmax over the L*P logits + sampling locations (ref * valid_ratio + offset / (W,H) | offset / P * wh / 2)
    + the bilinear gather; value (B,S,M,32) may be a column block of a wider matrix"""
    _count("msda_fused")
    B, S, M, D = value.shape
    geo = _geo(shapes_host, n_levels, B, S)
    assert [int(t) for t in lsi_host][:n_levels] == [sum(h * w for h, w in geo["level_hw"][:l]) for l in range(n_levels)]
    loc, attn = tdb.msda_prep(proj.float(), ref, valid_ratios, geo, Lq, M, P)
    return tdb.msda_core(value.float(), loc, attn, geo["level_hw"]).to(value.dtype).contiguous()


# ---- the fused kernels of the 16-bit throughput modes (operands and stored activations rounded to 16 bits, fp32 accumulation)
def gemm_ln(a, w, bias, residual, gamma, beta, add2=None, eps=1e-5, out=None, out2=None):
    """dtlr_gemm_ln: y = LN_256(a W^T + bias (+ residual)) [, y2 = y + add2]"""
    _count("gemm_ln")
    assert a.dtype in HALF and w.dtype == a.dtype and w.shape[0] == 256
    z = a.float() @ w.float().t() + bias
    if residual is not None:
        z = z + residual.float()
    y = F.layer_norm(z, (256,), gamma, beta, eps).to(a.dtype)
    return (y, (y.float() + add2.float()).to(a.dtype)) if add2 is not None else y


def ffn_ln_half(x, w1, b1, w2, b2, gamma, beta, eps=1e-5, add2=None):
    """dtlr_ffn_ln_ws: y = LN(x + W2 relu(W1 x + b1) + b2) [, y2 = y + add2]; the hidden activation is a 16-bit MMA operand in tensor memory"""
    _count("ffn_fused")
    assert x.dtype in HALF and x.shape[1] == 256 and w1.shape[0] % 128 == 0
    h = F.relu(x.float() @ w1.float().t() + b1).to(x.dtype)
    y = F.layer_norm(h.float() @ w2.float().t() + b2 + x.float(), (256,), gamma, beta, eps).to(x.dtype)
    return (y, (y.float() + add2.float()).to(x.dtype)) if add2 is not None else y


def mlp_head(x, l1, l2, w3_f32, b3, ref=None):
    """dtlr_mlp_head: 256 -> 256 -> 256 -> 4 box MLP [+ sigmoid(delta + inverse_sigmoid(ref))], fp32 result"""
    _count("mlp_head")
    assert x.dtype in HALF and w3_f32.dtype == torch.float32 and w3_f32.shape == (4, 256)
    h = F.relu(x.float() @ l1[0].float().t() + l1[1]).to(x.dtype)
    h = F.relu(h.float() @ l2[0].float().t() + l2[1])
    d = h @ w3_f32.t() + b3
    return d if ref is None else (d + inverse_sigmoid(ref)).sigmoid()


def install(monkeypatch, half=False):
    if half:
        monkeypatch.setattr(ops, "gemm_ln", gemm_ln)
        monkeypatch.setattr(ops, "ffn_ln", ffn_ln_half)
        monkeypatch.setattr(ops, "mlp_head", mlp_head)
    _install_common(monkeypatch)


def _install_common(monkeypatch):
    """route the engine's leaf launches to the stand-ins above (pytest monkeypatch: undone at the end of the test)"""
    CALLS.clear()
    for name in ("split_cast", "gemm", "im2col", "conv2d_nhwc", "stem_conv", "maxpool3x3s2", "groupnorm_into", "pos_sine_into", "add_layernorm",
                 "add", "zero_masked_rows_", "enc_ref_points", "encoder_proposals", "rowmax", "sine_embed", "box_refine", "cast",
                 "mha_self_attention", "topk_select", "select_gather"):
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(msda_mod, "msda_forward_fused", msda_forward_fused)
    monkeypatch.setattr(L, "require_cuda", lambda *t: None)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())

"""Kernel namespace of the native fine-tune step (dtlr_b200/train_engine.py).

Every function is a thin wrapper (torch tensors in / out) over the C ABI of include/dtlr_b200.h -- csrc/train.cu for the backward
kernels, the forward kernels of the inference engine for everything else.  The training engine only ever talks to this namespace,
so its chain-rule orchestration can be checked on the CPU against autograd with a torch stand-in for this module living under
tests/ (tests/train_ops_double.py); the product has no such fallback: without the CUDA library these calls raise.
"""
import ctypes

import torch
import torch.nn.functional as F

from . import _lib as L
from . import msda as msda_mod
from . import ops

HALF = ops.HALF
_call, _p, _st = ops._call, ops._p, ops._st


def check_device(*tensors):
    L.require_cuda(*tensors)


# ------------------------------------------------------------------------------------------------------------ forward pieces
def gemm(a, w, bias=None, residual=None, relu=0, out_dtype=None, out=None):
    return ops.gemm(a, w, bias, residual=residual, relu=relu, out_dtype=out_dtype, out=out)


def layernorm(z, gamma, beta, add2=None):
    """y = LN(z) [, y + add2]"""
    return ops.add_layernorm(z, None, gamma, beta, add2=add2)


def add(a, b):
    return ops.add(a, b)


def cast(x, dtype):
    return ops.cast(x.contiguous(), dtype)


def zero_masked_rows_(x, pad_u8):
    return ops.zero_masked_rows_(x, pad_u8)


def enc_ref_points(vr, geo):
    return ops.enc_ref_points(vr, geo["shapes_host"], geo["nlev"], geo["B"], geo["S"])


def msda_prep(oa_f32, ref, vr, geo, Lq, M, P):
    return ops.msda_prep(oa_f32, ref, vr, geo["shapes_host"], geo["nlev"], geo["B"], Lq, M, P)


def msda_forward(val4, loc, attn, geo):
    """val4 fp32 (B,S,M,D) contiguous -> core fp32 (B*Lq, M*D)"""
    out = msda_mod.msda_forward_raw(val4, geo["shapes_host"], geo["lsi_host"], geo["nlev"], loc, attn)
    return out.view(-1, out.shape[-1])


def msda_backward(val4, loc, attn, gout, geo):
    """-> grad_value (B,S,M,D), grad_loc, grad_attn (fp32)"""
    B, S, M, D = val4.shape
    Lq, P = loc.shape[1], loc.shape[4]
    gout = gout.contiguous()
    gv = torch.empty_like(val4)
    gl = torch.empty_like(loc)
    ga = torch.empty_like(attn)
    with torch.cuda.device(val4.device):
        rc = L.lib().dtlr_msda_backward(_p(val4), geo["shapes_host"], geo["lsi_host"], _p(loc), _p(attn), _p(gout), _p(gv), _p(gl), _p(ga),
                                        B, S, M, D, geo["nlev"], Lq, P, L.dtype_code(val4), _st(val4))
    L.check(rc, "dtlr_msda_backward")
    return gv, gl, ga


def msda_bwd_glue(gl, ga, attn, ref, vr, geo, Lq, M, P, out_dtype):
    B = geo["B"]
    n = M * geo["nlev"] * P * 3
    out = torch.empty((B * Lq, n), dtype=out_dtype, device=gl.device)
    _call("dtlr_msda_bwd_glue", _p(gl), _p(ga), _p(attn), _p(ref), ref.shape[-1], _p(vr), geo["shapes_host"], geo["nlev"], _p(out),
          out.stride(0), B, Lq, M, P, L.dtype_code(out), _st(gl))
    return out


def sine_embed(ref, vr, B, Q, nlev, dtype):
    return ops.sine_embed(ref, vr, B, Q, nlev, dtype)


# ------------------------------------------------------------------------------------------------------------ backward kernels
def wgrad(dy, x, gw):
    """gw[n, k] += sum_r dy[r, n] x[r, k]; dy (rows, N) / x (rows, K) of one dtype (16-bit or fp32), gw fp32 (N, K) view of the arena"""
    L.require_cuda(dy, x, gw)
    assert dy.dim() == 2 and x.dim() == 2 and dy.shape[0] == x.shape[0] and dy.stride(1) == 1 and x.stride(1) == 1 and dy.dtype == x.dtype
    assert gw.dtype == torch.float32 and gw.shape == (dy.shape[1], x.shape[1]) and gw.stride(1) == 1
    _call("dtlr_wgrad", _p(dy), dy.stride(0), _p(x), x.stride(0), _p(gw), gw.stride(0), dy.shape[0], dy.shape[1], x.shape[1],
          L.dtype_code(dy), _st(dy))


def colsum(x, out, nseg=1, seg_rows=None, seg_stride=0, row0=0):
    """out[c] += sum of x[rows, c] over nseg segments of seg_rows rows (segment s starts at row row0 + s*seg_stride); out fp32 (N)"""
    L.require_cuda(x, out)
    assert x.dim() == 2 and x.stride(1) == 1 and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == x.shape[1]
    if seg_rows is None:
        seg_rows = x.shape[0]
    xs = x[row0:] if row0 else x
    _call("dtlr_colsum", _p(xs), ctypes.c_longlong(x.stride(0)), x.shape[1], ctypes.c_longlong(nseg), ctypes.c_longlong(seg_rows),
          ctypes.c_longlong(seg_stride), _p(out), L.dtype_code(x), _st(x))


def layernorm_bwd(z, dy, dy2, gamma, dgamma, dbeta, want32=True, want16=True, eps=1e-5):
    """LayerNorm(256) backward from the saved pre-norm rows z; dy (+ dy2) fp32.  Returns (dz fp32 | None, dz of z.dtype | None);
    dgamma / dbeta (fp32 views of the gradient arena, or None) are accumulated."""
    L.require_cuda(z, dy)
    rows, C = z.shape
    assert z.is_contiguous() and dy.is_contiguous() and dy.dtype == torch.float32 and dy.shape == z.shape
    assert dy2 is None or (dy2.is_contiguous() and dy2.dtype == torch.float32 and dy2.shape == z.shape)
    if z.dtype == torch.float32:        # one stream serves both uses
        want32, want16 = True, False
    dz32 = torch.empty((rows, C), dtype=torch.float32, device=z.device) if want32 else None
    dz16 = torch.empty((rows, C), dtype=z.dtype, device=z.device) if want16 else None
    _call("dtlr_layernorm_bwd", _p(z), _p(dy), _p(dy2), _p(gamma), _p(dz32), _p(dz16), _p(dgamma), _p(dbeta), ctypes.c_longlong(rows), C,
          ctypes.c_float(eps), L.dtype_code(z), _st(z))
    if z.dtype == torch.float32:
        return dz32, dz32
    return dz32, dz16


def relu_bwd_(dh, h):
    assert dh.is_contiguous() and h.is_contiguous() and dh.shape == h.shape and dh.dtype == h.dtype
    _call("dtlr_relu_bwd", _p(dh), _p(h), ctypes.c_longlong(dh.numel()), L.dtype_code(dh), _st(dh))
    return dh


def add_cast(a, b, c, out_dtype):
    """a (+ b) (+ c): fp32 inputs -> out_dtype"""
    assert a.dtype == torch.float32 and a.is_contiguous() and (b is None or (b.is_contiguous() and b.dtype == torch.float32))
    assert c is None or (c.is_contiguous() and c.dtype == torch.float32)
    out = torch.empty(a.shape, dtype=out_dtype, device=a.device)
    _call("dtlr_add_cast", _p(a), _p(b), _p(c), _p(out), ctypes.c_longlong(a.numel()), L.dtype_code(out), _st(a))
    return out


def pack_weights(table_dev, tile_start_dev, n_entries, total_tiles, dtype):
    L.set_flavor(dtype)
    code = L._DT[dtype]
    _call("dtlr_pack_weights", _p(table_dev), _p(tile_start_dev), int(n_entries), int(total_tiles), code, _st(table_dev))


def optim_begin(state):
    _call("dtlr_optim_begin", _p(state), _st(state))


def grad_sumsq(g, state):
    _call("dtlr_grad_sumsq", _p(g), ctypes.c_longlong(g.numel()), _p(state), _st(g))


def adamw(p, g, m, v, lr, beta1, beta2, eps, weight_decay, max_norm, state):
    _call("dtlr_adamw", _p(p), _p(g), _p(m), _p(v), ctypes.c_longlong(p.numel()), ctypes.c_float(lr), ctypes.c_float(beta1),
          ctypes.c_float(beta2), ctypes.c_float(eps), ctypes.c_float(weight_decay), ctypes.c_float(max_norm), _p(state), _st(p))


# ------------------------------------------------------------------------------------------------------------ decoder self-attention
def make_mask(mask_bool):
    """(Q,Q) bool, True = blocked (dn_components.py:121-141) -> the bit matrices of csrc/attention_train.cu in both orientations
    (uint32 words stored as int32: word w of row i holds keys 32w .. 32w+31)"""
    if mask_bool is None:
        return None
    Q = mask_bool.shape[0]
    KP = (Q + 63) // 64 * 64
    sh = torch.arange(32, device=mask_bool.device, dtype=torch.int64)

    def bits(m):
        full = torch.zeros((Q, KP), dtype=torch.int64, device=m.device)
        full[:, :Q] = m
        w = (full.view(Q, KP // 32, 32) << sh).sum(-1)
        return torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32).contiguous()
    return {"bool": mask_bool, "bits": bits(mask_bool), "bitsT": bits(mask_bool.t())}


def sa_forward(qk, v, mask, B, Q, heads):
    """nn.MultiheadAttention core (deformable_transformer.py:847, 903-905) with what the backward needs.  qk (B*Q, 2d) = [q | k]
    projections, v (B*Q, d); mask = make_mask(...) or None.  16-bit: the flash kernels of csrc/attention_train.cu (forward keeps only
    the per-row log-sum-exp); fp32 parity mode: torch scaled_dot_product_attention under autograd.  Returns (att (B*Q, d), ctx)."""
    d = v.shape[1]
    hd = d // heads
    if qk.dtype in HALF and hd == 32:
        out = torch.empty((B * Q, d), dtype=v.dtype, device=v.device)
        lse2 = torch.empty((B, heads, Q), dtype=torch.float32, device=v.device)
        bits = mask["bits"] if mask is not None else None
        _call("dtlr_mha_train_forward", _p(qk), qk.stride(0), d, _p(v), v.stride(0), _p(bits), _p(out), out.stride(0), _p(lse2), B, Q, heads,
              hd, L.dtype_code(v), _st(v))
        return out, ("native", qk, v, out, lse2, mask, B, Q, heads, hd)
    q4 = qk[:, :d].reshape(B, Q, heads, hd).transpose(1, 2).detach().requires_grad_(True)
    k4 = qk[:, d:].reshape(B, Q, heads, hd).transpose(1, 2).detach().requires_grad_(True)
    v4 = v.reshape(B, Q, heads, hd).transpose(1, 2).detach().requires_grad_(True)
    allow = None if mask is None else ~(mask["bool"] if isinstance(mask, dict) else mask)
    with torch.enable_grad():
        o = F.scaled_dot_product_attention(q4, k4, v4, attn_mask=allow)
    att = o.detach().transpose(1, 2).reshape(B * Q, d)
    return att, ("sdpa", o, q4, k4, v4, B, Q, heads, hd)


def sa_backward(ctx, datt):
    """datt (B*Q, d) -> (dqk (B*Q, 2d), dv (B*Q, d)) of datt.dtype"""
    if ctx[0] == "native":
        _, qk, v, out, lse2, mask, B, Q, heads, hd = ctx
        d = heads * hd
        datt = datt.contiguous()
        dqk = torch.empty((B * Q, 2 * d), dtype=qk.dtype, device=qk.device)
        dv = torch.empty((B * Q, d), dtype=qk.dtype, device=qk.device)
        Dv = torch.empty((B, heads, Q), dtype=torch.float32, device=qk.device)
        bits = mask["bits"] if mask is not None else None
        bitsT = mask["bitsT"] if mask is not None else None
        _call("dtlr_mha_train_backward", _p(qk), qk.stride(0), d, _p(v), v.stride(0), _p(out), out.stride(0), _p(datt), datt.stride(0),
              _p(bits), _p(bitsT), _p(lse2), _p(Dv), _p(dqk), dqk.stride(0), _p(dv), dv.stride(0), B, Q, heads, hd, L.dtype_code(v), _st(v))
        L.LAUNCHES += 2
        return dqk, dv
    _, o, q4, k4, v4, B, Q, heads, hd = ctx
    g = datt.reshape(B, Q, heads, hd).transpose(1, 2).to(o.dtype)
    dq, dk, dv = torch.autograd.grad(o, (q4, k4, v4), g)
    d = heads * hd
    dqk = torch.cat([dq.transpose(1, 2).reshape(B * Q, d), dk.transpose(1, 2).reshape(B * Q, d)], 1)
    return dqk, dv.transpose(1, 2).reshape(B * Q, d).contiguous()


# ------------------------------------------------------------------------------------------------------------ gradient-free pieces
def two_stage_refs(eng, memory, geo):
    """Two-stage query selection of deformable_transformer.py:320-353 on the inference kernels (no gradient reaches it in the CTC
    fine-tune step: the anchors are detached and tgt is the learned table).  Returns sigmoid(refpoint_embed) fp32 (B, Q, 4)."""
    from .engine import InferenceEngine as IE
    tr = eng.model.transformer
    B, S, d = geo["B"], geo["S"], tr.d_model
    om, prop = ops.encoder_proposals(memory, geo["pad_u8"], geo["valid_hw"], geo["shapes_host"], geo["nlev"], B, S, d, tr.two_stage_default_hw)
    W = eng.lin
    omn = layernorm(gemm(om, W["enc_output"].w16, W["enc_output"].bias), *eng.ln_params(tr.enc_output_norm))
    cls = IE._head_gemm(omn, W["enc_cls"].w16, W["enc_cls"].bias)
    scores = ops.rowmax(cls, cls.shape[1]).view(B, S)
    bb = W["enc_bbox"]
    delta = box_head(eng, omn, bb, None).view(B, S, 4)
    topk = ops.topk_select(scores, tr.num_queries)
    if tr.debug_force_topk is not None:
        topk = tr.debug_force_topk.to(memory.device).contiguous()
    ref0, _, _ = ops.select_gather(topk, delta, prop.view(B, S, 4), omn.view(B, S, d))
    return ref0


def box_head(eng, x, lins, ref):
    """bbox MLP (models/dino/utils.py:110-122) [+ refinement against ref (deformable_transformer.py:734-738)], no gradient"""
    l0, l1, l2 = lins
    if x.dtype in HALF and ops.MLP_HEAD_FUSED and x.stride(0) % 8 == 0:
        return ops.mlp_head(x, (l0.w16, l0.bias), (l1.w16, l1.bias), l2.master_w.contiguous(), l2.bias, ref)
    h = gemm(gemm(x, l0.w16, l0.bias, relu=1), l1.w16, l1.bias, relu=1)
    delta = gemm(h, l2.w16, l2.bias, out_dtype=torch.float32)
    return delta if ref is None else ops.box_refine(delta, ref)


def head_logits(x, lin):
    """fp32 class logits with a 16-byte row pitch (a column view of the padded buffer)"""
    from .engine import InferenceEngine as IE
    return IE._head_gemm(x, lin.w16, lin.bias)


def ctc_loss_grad(logits, boxes, targets_i32, lens_i32, eps=0.003, zero_infinity=True):
    """loss_CTC (models/dino/dino.py:457-551) and d loss / d logits from the fused kernel set (csrc/decode.cu: dtlr_ctc_loss).
    logits fp32 (B,Q,C) (rows may be pitched), boxes fp32 (B,Q,4).  Returns (loss 0-d tensor, grad fp32 (B,Q,C) contiguous)."""
    B, Q, C = logits.shape
    assert logits.dtype == torch.float32 and logits.stride(2) == 1 and logits.stride(0) == Q * logits.stride(1)
    boxes = boxes.float().contiguous()
    dev = logits.device
    Lmax = int(targets_i32.shape[1])
    S = 2 * Lmax + 1
    nll = torch.empty((B,), dtype=torch.float32, device=dev)
    grad = torch.empty((B, Q, C), dtype=torch.float32, device=dev)
    perm = torch.empty((B, Q), dtype=torch.int32, device=dev)
    rsum = torch.empty((B, Q), dtype=torch.float32, device=dev)
    label = torch.empty((B, Q), dtype=torch.int32, device=dev)
    frames = torch.empty((B, Q), dtype=torch.int32, device=dev)
    lp = torch.empty((B, Q, Lmax + 1), dtype=torch.float32, device=dev)
    alpha = torch.empty((B, Q, S), dtype=torch.float32, device=dev)
    gext = torch.empty((B, Q, S), dtype=torch.float32, device=dev)
    _call("dtlr_ctc_loss", _p(logits), logits.stride(1), _p(boxes), _p(targets_i32), _p(lens_i32), Lmax, ctypes.c_float(eps),
          1 if zero_infinity else 0, _p(nll), _p(grad), _p(perm), _p(rsum), _p(label), _p(frames), _p(lp), _p(alpha), _p(gext),
          B, Q, C, _st(logits))
    L.LAUNCHES += 4
    per = nll / lens_i32.clamp(min=1).float()
    if zero_infinity:
        per = torch.where(torch.isinf(nll), torch.zeros_like(per), per)
    return per.mean(), grad

"""TEST INFRASTRUCTURE -- a plain-torch (CPU, fp32) stand-in for dtlr_b200/train_ops.py.

It exists so that the chain-rule orchestration of dtlr_b200/train_engine.py (which launch consumes which saved tensor, which gradient
is added where) can be checked on the CPU against torch autograd of the module path.  It is NOT a fallback: nothing under dtlr_b200/
imports it; each function states the formula the corresponding CUDA kernel implements (csrc/train.cu, csrc/nn_ops.cu, csrc/msda.cu).
"""
import math

import torch
import torch.nn.functional as F

from dtlr_b200.train_ops import sa_backward, sa_forward          # pure torch (scaled_dot_product_attention), device agnostic
from dtlr_b200.transformer import gen_encoder_output_proposals, gen_sineembed_for_position
from dtlr_b200.misc import inverse_sigmoid


def check_device(*tensors):
    pass


def make_mask(mask_bool):
    return mask_bool


# ------------------------------------------------------------------------------------------------------------ forward pieces
def gemm(a, w, bias=None, residual=None, relu=0, out_dtype=None, out=None):
    c = a.float() @ w.float().t()
    if bias is not None:
        c = c + bias
    if relu == 1:
        c = F.relu(c)
    if residual is not None:
        c = torch.where(residual.float() > 0, c, torch.zeros_like(c)) if relu == 3 else c + residual.float()
    if relu == 2:
        c = F.relu(c)
    c = c.to(out_dtype or a.dtype)
    if out is not None:
        out.copy_(c)
        return out
    return c


def layernorm(z, gamma, beta, add2=None):
    y = F.layer_norm(z.float(), (z.shape[-1],), gamma, beta, 1e-5).to(z.dtype)
    return (y, (y.float() + add2.float()).to(z.dtype)) if add2 is not None else y


def add(a, b):
    return a + b


def cast(x, dtype):
    return x.to(dtype)


def zero_masked_rows_(x, pad_u8):
    x[pad_u8.bool()] = 0
    return x


def enc_ref_points(vr, geo):
    B, S = geo["B"], geo["S"]
    out = []
    for l, (H, W) in enumerate(geo["level_hw"]):
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
        rx = (xs.reshape(-1)[None] + 0.5) / (vr[:, l, 0:1] * W)
        ry = (ys.reshape(-1)[None] + 0.5) / (vr[:, l, 1:2] * H)
        out.append(torch.stack([rx, ry], -1))
    return torch.cat(out, 1).reshape(B * S, 2).contiguous()


def _scales(ref, vr, geo, Lq, P):
    """d loc / d offset per (b, q, level): (sx, sy), each (B, Lq, L)"""
    B, L = geo["B"], geo["nlev"]
    if ref.shape[-1] == 2:
        sx = torch.tensor([1.0 / w for h, w in geo["level_hw"]]).view(1, 1, L).expand(B, Lq, L)
        sy = torch.tensor([1.0 / h for h, w in geo["level_hw"]]).view(1, 1, L).expand(B, Lq, L)
    else:
        r = ref.view(B, Lq, 4)
        sx = r[:, :, 2:3] * vr[:, None, :, 0] * 0.5 / P
        sy = r[:, :, 3:4] * vr[:, None, :, 1] * 0.5 / P
    return sx, sy


def msda_prep(oa, ref, vr, geo, Lq, M, P):
    B, L = geo["B"], geo["nlev"]
    n = M * L * P
    off = oa[:, :2 * n].reshape(B, Lq, M, L, P, 2)
    attn = F.softmax(oa[:, 2 * n:3 * n].reshape(B, Lq, M, L * P), -1).reshape(B, Lq, M, L, P)
    r = ref.view(B, Lq, -1)
    cx = r[:, :, 0:1] * vr[:, None, :, 0]           # (B,Lq,L)
    cy = r[:, :, 1:2] * vr[:, None, :, 1]
    sx, sy = _scales(ref, vr, geo, Lq, P)
    x = cx[:, :, None, :, None] + off[..., 0] * sx[:, :, None, :, None]
    y = cy[:, :, None, :, None] + off[..., 1] * sy[:, :, None, :, None]
    return torch.stack([x, y], -1).contiguous(), attn.contiguous()


def msda_core(val4, loc, attn, level_hw):
    """out[b,q,m,:] = sum_{l,p} attn * bilinear(value_l, loc) with pixel = loc * size - 0.5 and zero padding (differentiable)"""
    B, S, M, D = val4.shape
    _, Lq, _, L, P, _ = loc.shape
    out = 0
    s0 = 0
    for l, (H, W) in enumerate(level_hw):
        v = val4[:, s0:s0 + H * W].permute(0, 2, 3, 1).reshape(B * M, D, H, W)
        s0 += H * W
        g = (2 * loc[:, :, :, l] - 1).permute(0, 2, 1, 3, 4).reshape(B * M, Lq, P, 2)
        smp = F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False)      # (B*M, D, Lq, P)
        w = attn[:, :, :, l].permute(0, 2, 1, 3).reshape(B * M, 1, Lq, P)
        out = out + (smp * w).sum(-1)
    return out.view(B, M, D, Lq).permute(0, 3, 1, 2).reshape(B, Lq, M * D)


def msda_forward(val4, loc, attn, geo):
    out = msda_core(val4, loc, attn, geo["level_hw"])
    return out.reshape(-1, out.shape[-1])


def msda_backward(val4, loc, attn, gout, geo):
    v, l, a = (t.detach().clone().requires_grad_(True) for t in (val4, loc, attn))
    with torch.enable_grad():
        out = msda_core(v, l, a, geo["level_hw"])
    return torch.autograd.grad(out, (v, l, a), gout.reshape(out.shape))


def msda_bwd_glue(gl, ga, attn, ref, vr, geo, Lq, M, P, out_dtype):
    B, L = geo["B"], geo["nlev"]
    sx, sy = _scales(ref, vr, geo, Lq, P)
    doff = torch.stack([gl[..., 0] * sx[:, :, None, :, None], gl[..., 1] * sy[:, :, None, :, None]], -1)
    a = attn.reshape(B, Lq, M, L * P)
    g = ga.reshape(B, Lq, M, L * P)
    dlog = a * (g - (a * g).sum(-1, keepdim=True))
    return torch.cat([doff.reshape(B * Lq, -1), dlog.reshape(B * Lq, -1)], 1).to(out_dtype)


def sine_embed(ref, vr, B, Q, nlev, dtype):
    r = ref.view(B, Q, 4) * torch.cat([vr[:, 0], vr[:, 0]], -1)[:, None]
    return gen_sineembed_for_position(r).reshape(B * Q, -1).to(dtype)


# ------------------------------------------------------------------------------------------------------------ backward kernels
def wgrad(dy, x, gw):
    gw += dy.float().t() @ x.float()


def colsum(x, out, nseg=1, seg_rows=None, seg_stride=0, row0=0):
    if seg_rows is None:
        seg_rows = x.shape[0]
    for s in range(nseg):
        a = row0 + s * seg_stride
        out += x[a:a + seg_rows].float().sum(0)


def layernorm_bwd(z, dy, dy2, gamma, dgamma, dbeta, want32=True, want16=True, eps=1e-5):
    d = dy if dy2 is None else dy + dy2
    x = z.float()
    mean = x.mean(-1, keepdim=True)
    xc = x - mean
    rstd = torch.rsqrt((xc * xc).mean(-1, keepdim=True) + eps)
    xh = xc * rstd
    g = d * gamma
    dz = rstd * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    if dgamma is not None:
        dgamma += (d * xh).sum(0)
    if dbeta is not None:
        dbeta += d.sum(0)
    return dz, dz.to(z.dtype)


def relu_bwd_(dh, h):
    dh[h <= 0] = 0
    return dh


def add_cast(a, b, c, out_dtype):
    r = a
    if b is not None:
        r = r + b
    if c is not None:
        r = r + c
    return r.to(out_dtype)


def repack_lins(lins, dtype):
    for l in lins:
        w = torch.cat([p[0].detach()[p[1]:p[1] + p[2]] for p in l.parts], 0)
        l.w16.copy_(w.to(dtype))
        if l.wT16 is not None:
            l.wT16[:, :l.N] = w.t().to(dtype)


def pack_weights(*a):
    raise AssertionError("the double repacks through repack_lins")


def optim_begin(state):
    state[0] = 0
    state[1] += 1


def grad_sumsq(g, state):
    state[0] += (g.double() ** 2).sum().float()


def adamw(p, g, m, v, lr, beta1, beta2, eps, weight_decay, max_norm, state):
    """torch.nn.utils.clip_grad_norm_ coefficient + torch.optim.AdamW single-tensor update"""
    norm = float(state[0]) ** 0.5
    clip = min(1.0, max_norm / (norm + 1e-6)) if max_norm > 0 else 1.0
    step = float(state[1])
    gi = g * clip
    p.mul_(1 - lr * weight_decay)
    m.mul_(beta1).add_(gi, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(gi, gi, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p.addcdiv_(m, denom, value=-lr / bc1)


# ------------------------------------------------------------------------------------------------------------ gradient-free pieces
def box_head(eng, x, lins, ref):
    h = x.float()
    for i, l in enumerate(lins):
        h = h @ l.w16.float().t() + l.bias
        if i < len(lins) - 1:
            h = F.relu(h)
    return h if ref is None else (h + inverse_sigmoid(ref)).sigmoid()


def head_logits(x, lin):
    return x.float() @ lin.w16.float().t() + lin.bias


def two_stage_refs(eng, memory, geo):
    tr = eng.model.transformer
    B, S, d = geo["B"], geo["S"], tr.d_model
    pad = geo["pad_u8"].bool().view(B, S)
    om, prop = gen_encoder_output_proposals(memory.float().view(B, S, d), pad, list(geo["level_hw"]), tr.two_stage_default_hw)
    W = eng.lin
    om = layernorm(gemm(om.reshape(B * S, d), W["enc_output"].w16, W["enc_output"].bias), *eng.ln_params(tr.enc_output_norm))
    cls = head_logits(om, W["enc_cls"])
    coord = box_head(eng, om, W["enc_bbox"], None).view(B, S, 4) + prop
    topk = torch.topk(cls.max(-1)[0].view(B, S), tr.num_queries, dim=1)[1]
    if tr.debug_force_topk is not None:
        topk = tr.debug_force_topk
    return torch.gather(coord, 1, topk.unsqueeze(-1).repeat(1, 1, 4)).sigmoid()


def ctc_loss_grad(logits, boxes, targets_i32, lens_i32, eps=0.003, zero_infinity=True):
    """the reference's literal torch chain (dtlr_b200.dino.SetCriterion.loss_CTC, CPU branch) under autograd"""
    from dtlr_b200.dino import SetCriterion
    crit = SetCriterion(logits.shape[-1], None, {}, 0.25, [])
    lg = logits.detach().clone().requires_grad_(True)
    tg = [{"labels": targets_i32[i, :int(lens_i32[i])].long()} for i in range(logits.shape[0])]
    with torch.enable_grad():
        loss = crit.loss_CTC({"pred_logits": lg, "pred_boxes": boxes}, tg, None, None)["loss_CTC"]
    (g,) = torch.autograd.grad(loss, lg)
    return loss.detach(), g.contiguous()

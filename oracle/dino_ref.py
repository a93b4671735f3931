"""ORACLE -- test infrastructure, NOT product code.

CPU restatement (plain torch CPU tensor arithmetic + the C MSDA core in oracle/msda_ref.c) of the DTLR
DINO-DETR forward path, written from a reading of the reference; every function cites the reference
file:line it follows.  It works on a plain {key: tensor} state_dict with the reference key layout and
has no nn.Module, no autograd requirement and no dependence on dtlr_b200's model code.

Pinning: the reference publishes golden vectors only for the MSDA core (models/dino/ops/test.py); for the
rest, tests/golden/make_golden.py runs the UNMODIFIED reference (imported from /root/reference under the
shims of tests/golden/ref_shims.py) on synthetic weights/inputs and commits its outputs under tests/golden/;
tests/test_oracle_dino.py checks this restatement against those vectors.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this module.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from . import msda as _msda


def default_cfg(**kw):
    """Hot-path keys of reference config/Latin_CTC.py:3-116."""
    c = dict(num_classes=166, hidden_dim=256, nheads=8, num_queries=900, enc_layers=6, dec_layers=6,
             dim_feedforward=2048, num_feature_levels=4, enc_n_points=4, dec_n_points=4,
             pe_temperatureH=20, pe_temperatureW=20, dn_number=0, two_stage_default_hw=0.05)
    c.update(kw)
    return SimpleNamespace(**c)


# ----------------------------------------------------------------------------------------------- helpers
def inverse_sigmoid(x, eps=1e-3):
    """reference util/misc.py:575-579"""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def _linear(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _ln(sd, prefix, x):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def _mlp(sd, prefix, x, n):
    """reference models/dino/utils.py:110-122"""
    for i in range(n):
        x = _linear(sd, "%s.layers.%d" % (prefix, i), x)
        if i < n - 1:
            x = F.relu(x)
    return x


def nested_from_list(tensors):
    """reference util/misc.py:375-397: zero-pad to the max H,W; mask True on padding."""
    if torch.is_tensor(tensors) and tensors.dim() == 4:
        tensors = list(tensors)
    C = tensors[0].shape[0]
    H = max(t.shape[1] for t in tensors)
    W = max(t.shape[2] for t in tensors)
    out = torch.zeros(len(tensors), C, H, W, dtype=tensors[0].dtype)
    mask = torch.ones(len(tensors), H, W, dtype=torch.bool)
    for i, t in enumerate(tensors):
        out[i, :, : t.shape[1], : t.shape[2]] = t
        mask[i, : t.shape[1], : t.shape[2]] = False
    return out, mask


def resize_mask(mask, size):
    """reference models/dino/backbone.py:103 / dino.py:304-307: nearest, src = floor(dst*in/out)."""
    return F.interpolate(mask[None].float(), size=size).to(torch.bool)[0]


# ----------------------------------------------------------------------------------------------- backbone
def _frozen_bn(sd, prefix, x):
    """reference models/dino/backbone.py:62-72 (eps inside the rsqrt)."""
    scale = sd[prefix + ".weight"] * (sd[prefix + ".running_var"] + 1e-5).rsqrt()
    bias = sd[prefix + ".bias"] - sd[prefix + ".running_mean"] * scale
    return x * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)


def _bottleneck(sd, p, x, stride):
    """torchvision ResNet v1.5 Bottleneck (stride on the 3x3), as instantiated by reference backbone.py:118-120."""
    idt = x
    y = F.relu(_frozen_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"])))
    y = F.relu(_frozen_bn(sd, p + ".bn2", F.conv2d(y, sd[p + ".conv2.weight"], stride=stride, padding=1)))
    y = _frozen_bn(sd, p + ".bn3", F.conv2d(y, sd[p + ".conv3.weight"]))
    if (p + ".downsample.0.weight") in sd:
        idt = _frozen_bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride))
    return F.relu(y + idt)


def resnet50_body(sd, x, prefix="backbone.0.body"):
    """reference backbone.py:83-106: IntermediateLayerGetter returning layer2/3/4 (return_interm_indices=[1,2,3])."""
    y = F.conv2d(x, sd[prefix + ".conv1.weight"], stride=2, padding=3)
    y = F.relu(_frozen_bn(sd, prefix + ".bn1", y))
    y = F.max_pool2d(y, kernel_size=3, stride=2, padding=1)
    feats = []
    for li, nblk in enumerate((3, 4, 6, 3), start=1):
        for b in range(nblk):
            stride = 2 if (b == 0 and li > 1) else 1
            y = _bottleneck(sd, "%s.layer%d.%d" % (prefix, li, b), y, stride)
        if li >= 2:
            feats.append(y)
    return feats


def pos_sine_hw(mask, temp_h, temp_w, npf):
    """reference models/dino/position_encoding.py:79-108 (normalize=True, scale=2*pi)."""
    not_mask = ~mask
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    eps = 1e-6
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * (2 * math.pi)
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * (2 * math.pi)
    i = torch.arange(npf, dtype=torch.float32)
    dim_tx = temp_w ** (2 * (i // 2) / npf)
    dim_ty = temp_h ** (2 * (i // 2) / npf)
    px = x_embed[:, :, :, None] / dim_tx
    py = y_embed[:, :, :, None] / dim_ty
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((py, px), dim=3).permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------------------------- MSDeformAttn
def msdeform_attn(sd, p, query, ref, src, shapes, lsi, pad_mask, M, L, P, st=None):
    """reference models/dino/ops/modules/ms_deform_attn.py:78-126"""
    N, Lq, C = query.shape
    S = src.shape[1]
    value = _linear(sd, p + ".value_proj", src)
    if pad_mask is not None:
        value = value.masked_fill(pad_mask[..., None], 0.0)
    value = value.view(N, S, M, C // M)
    off = _linear(sd, p + ".sampling_offsets", query).view(N, Lq, M, L, P, 2)
    aw = _linear(sd, p + ".attention_weights", query).view(N, Lq, M, L * P)
    aw = F.softmax(aw, -1).view(N, Lq, M, L, P)
    if ref.shape[-1] == 2:
        norm = torch.stack([shapes[..., 1], shapes[..., 0]], -1).to(query.dtype)
        loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = ref[:, :, None, :, None, :2] + off / P * ref[:, :, None, :, None, 2:] * 0.5
    out = _msda.msda_forward(value, shapes, lsi, loc, aw).to(query.dtype)
    if st is not None:
        st.update(value=value, loc=loc, aw=aw, core=out)
    return _linear(sd, p + ".output_proj", out)


# ----------------------------------------------------------------------------------------------- encoder
def encoder_reference_points(shapes, valid_ratios):
    """reference deformable_transformer.py:479-492"""
    pts = []
    for lvl, (H, W) in enumerate(shapes.tolist()):
        ry, rx = torch.meshgrid(torch.linspace(0.5, H - 0.5, H), torch.linspace(0.5, W - 0.5, W), indexing="ij")
        ry = ry.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H)
        rx = rx.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W)
        pts.append(torch.stack((rx, ry), -1))
    ref = torch.cat(pts, 1)
    return ref[:, :, None] * valid_ratios[:, None]


def encoder_layer(sd, p, src, pos, ref, shapes, lsi, pad_mask, cfg, st=None):
    """reference deformable_transformer.py:804-823 (post-norm, dropout 0)"""
    a = msdeform_attn(sd, p + ".self_attn", src + pos, ref, src, shapes, lsi, pad_mask,
                      cfg.nheads, cfg.num_feature_levels, cfg.enc_n_points, st)
    src = _ln(sd, p + ".norm1", src + a)
    f = _linear(sd, p + ".linear2", F.relu(_linear(sd, p + ".linear1", src)))
    return _ln(sd, p + ".norm2", src + f)


def gen_encoder_output_proposals(memory, pad_mask, shapes, default_hw=0.05):
    """reference models/dino/utils.py:15-64"""
    N, S, C = memory.shape
    props = []
    cur = 0
    for lvl, (H, W) in enumerate(shapes.tolist()):
        m = pad_mask[:, cur:cur + H * W].view(N, H, W)
        valid_h = (~m[:, :, 0]).sum(1)
        valid_w = (~m[:, 0, :]).sum(1)
        gy, gx = torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing="ij")
        grid = torch.stack([gx, gy], -1)
        scale = torch.stack([valid_w, valid_h], 1).view(N, 1, 1, 2)
        grid = (grid[None].expand(N, -1, -1, -1) + 0.5) / scale
        wh = torch.ones_like(grid) * default_hw * (2.0 ** lvl)
        props.append(torch.cat((grid, wh), -1).view(N, -1, 4))
        cur += H * W
    prop = torch.cat(props, 1)
    valid = ((prop > 0.01) & (prop < 0.99)).all(-1, keepdim=True)
    prop = torch.log(prop / (1 - prop))
    prop = prop.masked_fill(pad_mask.unsqueeze(-1), float("inf"))
    prop = prop.masked_fill(~valid, float("inf"))
    mem = memory.masked_fill(pad_mask.unsqueeze(-1), 0.0).masked_fill(~valid, 0.0)
    return mem, prop


# ----------------------------------------------------------------------------------------------- decoder
def gen_sineembed_for_position(pos):
    """reference models/dino/utils.py:141-167: (…,4) -> (…,512) in order (y, x, w, h), T=10000, scale 2*pi."""
    scale = 2 * math.pi
    dim_t = torch.arange(128, dtype=torch.float32)
    dim_t = 10000 ** (2 * (dim_t // 2) / 128)

    def emb(v):
        e = (v * scale)[..., None] / dim_t
        return torch.stack((e[..., 0::2].sin(), e[..., 1::2].cos()), dim=-1).flatten(-2)

    return torch.cat((emb(pos[..., 1]), emb(pos[..., 0]), emb(pos[..., 2]), emb(pos[..., 3])), dim=-1)


def mha_self_attention(sd, p, qk_in, v_in, nheads, attn_mask=None):
    """torch.nn.MultiheadAttention(256, 8) as used at reference deformable_transformer.py:847, 903-905
    (batch-first restatement: inputs (B,Q,C)); attn_mask bool, True = blocked."""
    B, Q, C = qk_in.shape
    Wi, bi = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(qk_in, Wi[:C], bi[:C])
    k = F.linear(qk_in, Wi[C:2 * C], bi[C:2 * C])
    v = F.linear(v_in, Wi[2 * C:], bi[2 * C:])
    dh = C // nheads
    q = q.view(B, Q, nheads, dh).transpose(1, 2)
    k = k.view(B, Q, nheads, dh).transpose(1, 2)
    v = v.view(B, Q, nheads, dh).transpose(1, 2)
    s = (q * (1.0 / math.sqrt(dh))) @ k.transpose(-1, -2)
    if attn_mask is not None:
        s = s.masked_fill(attn_mask[None, None], float("-inf"))
    o = (F.softmax(s, -1) @ v).transpose(1, 2).reshape(B, Q, C)
    return _linear(sd, p + ".out_proj", o)


def decoder_layer(sd, p, tgt, query_pos, ref_in, memory, shapes, lsi, pad_mask, cfg, attn_mask=None, st=None):
    """reference deformable_transformer.py:876-997 with module_seq ['sa','ca','ffn'] (batch-first restatement)."""
    qk = tgt + query_pos
    tgt = _ln(sd, p + ".norm2", tgt + mha_self_attention(sd, p + ".self_attn", qk, tgt, cfg.nheads, attn_mask))
    ca = msdeform_attn(sd, p + ".cross_attn", tgt + query_pos, ref_in, memory, shapes, lsi, pad_mask,
                       cfg.nheads, cfg.num_feature_levels, cfg.dec_n_points, st)
    tgt = _ln(sd, p + ".norm1", tgt + ca)
    f = _linear(sd, p + ".linear2", F.relu(_linear(sd, p + ".linear1", tgt)))
    return _ln(sd, p + ".norm3", tgt + f)


# ----------------------------------------------------------------------------------------------- DN (training-mode quirk Q3)
def prepare_for_cdn_nonoise(targets, sd, cfg):
    """reference models/dino/dn_components.py:20-142 in the only mode the DTLR CTC fine-tuning uses:
    dn_number=0 (use_dn=False) forced to one group, label_noise and box_noise applied to ZERO boxes.
    With boxes == 0 (reference datasets/IAM.py:100-102) the box noise term is multiplied by diff == 0, and the
    label noise only draws random numbers it never applies (chosen_indice is unused), so the result is
    deterministic: positive+negative copies of the GT label embeddings with inverse_sigmoid(0) boxes."""
    B = len(targets)
    known_num = [int(t["labels"].numel()) for t in targets]
    single_pad = max(known_num)
    dn_number = 1
    pad_size = single_pad * 2 * dn_number
    C = cfg.hidden_dim
    q_label = torch.zeros(B, pad_size, C)
    q_bbox = torch.zeros(B, pad_size, 4)
    label_enc = sd["label_enc.weight"]
    for b, t in enumerate(targets):
        n = known_num[b]
        emb = label_enc[t["labels"].long()]
        box = inverse_sigmoid(t["boxes"].float())
        for rep in range(2 * dn_number):
            q_label[b, single_pad * rep: single_pad * rep + n] = emb
            q_bbox[b, single_pad * rep: single_pad * rep + n] = box
    tgt_size = pad_size + cfg.num_queries
    attn_mask = torch.zeros(tgt_size, tgt_size, dtype=torch.bool)
    attn_mask[pad_size:, :pad_size] = True
    # dn_number == 1: both the i==0 and i==dn_number-1 branches fire with empty slices -> nothing else masked
    return q_label, q_bbox, attn_mask, {"pad_size": pad_size, "num_dn_group": dn_number}


# ----------------------------------------------------------------------------------------------- full forward
def dino_forward(sd, cfg, images, mask=None, targets=None, training=False, stages=None, force_topk=None):
    """reference models/dino/dino.py:270-415 (+ deformable_transformer.py:257-429, 652-766).
    images: (B,3,H,W) fp32 or list of (3,h,w).  Returns the reference's output dict (aux_outputs included)."""
    with torch.no_grad():
        if mask is None:
            images, mask = nested_from_list(images)
        st = stages if stages is not None else {}
        d, L = cfg.hidden_dim, cfg.num_feature_levels
        feats = resnet50_body(sd, images)
        srcs, masks, poss = [], [], []
        for l, f in enumerate(feats):
            m = resize_mask(mask, f.shape[-2:])
            y = F.conv2d(f, sd["input_proj.%d.0.weight" % l], sd["input_proj.%d.0.bias" % l])
            y = F.group_norm(y, 32, sd["input_proj.%d.1.weight" % l], sd["input_proj.%d.1.bias" % l], 1e-5)
            srcs.append(y); masks.append(m)
            poss.append(pos_sine_hw(m, cfg.pe_temperatureH, cfg.pe_temperatureW, d // 2))
        for l in range(len(feats), L):
            x_in = feats[-1] if l == len(feats) else srcs[-1]
            y = F.conv2d(x_in, sd["input_proj.%d.0.weight" % l], sd["input_proj.%d.0.bias" % l], stride=2, padding=1)
            y = F.group_norm(y, 32, sd["input_proj.%d.1.weight" % l], sd["input_proj.%d.1.bias" % l], 1e-5)
            m = resize_mask(mask, y.shape[-2:])
            srcs.append(y); masks.append(m)
            poss.append(pos_sine_hw(m, cfg.pe_temperatureH, cfg.pe_temperatureW, d // 2))
        st["feats"] = feats
        st["srcs"] = srcs

        # ---- transformer prologue (deformable_transformer.py:269-293)
        shapes = torch.tensor([list(s.shape[-2:]) for s in srcs], dtype=torch.long)
        lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
        src = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
        pad = torch.cat([m.flatten(1) for m in masks], 1)
        pos = torch.cat([p.flatten(2).transpose(1, 2) + sd["transformer.level_embed"][l].view(1, 1, -1)
                         for l, p in enumerate(poss)], 1)
        vr = []
        for m in masks:
            _, H, W = m.shape
            vh = (~m[:, :, 0]).sum(1).float() / H
            vw = (~m[:, 0, :]).sum(1).float() / W
            vr.append(torch.stack([vw, vh], -1))
        valid_ratios = torch.stack(vr, 1)
        st.update(src_flatten=src, pos=pos, pad=pad, shapes=shapes, lsi=lsi, valid_ratios=valid_ratios)

        # ---- encoder
        ref = encoder_reference_points(shapes, valid_ratios)
        out = src
        for i in range(cfg.enc_layers):
            lst = {} if i == 0 else None
            out = encoder_layer(sd, "transformer.encoder.layers.%d" % i, out, pos, ref, shapes, lsi, pad, cfg, lst)
            if lst is not None:
                st["enc0"] = lst
        memory = out
        st["memory"] = memory

        # ---- two-stage query selection (deformable_transformer.py:320-363)
        om, prop = gen_encoder_output_proposals(memory, pad, shapes, cfg.two_stage_default_hw)
        om = _ln(sd, "transformer.enc_output_norm", _linear(sd, "transformer.enc_output", om))
        cls_unsel = _linear(sd, "transformer.enc_out_class_embed", om)
        coord_unsel = _mlp(sd, "transformer.enc_out_bbox_embed", om, 3) + prop
        scores = cls_unsel.max(-1)[0]
        topk = torch.topk(scores, cfg.num_queries, dim=1)[1] if force_topk is None else force_topk
        st.update(topk_scores=scores, topk_idx=topk)
        refpoint_undetach = torch.gather(coord_unsel, 1, topk.unsqueeze(-1).repeat(1, 1, 4))
        init_box_proposal = torch.gather(prop, 1, topk.unsqueeze(-1).repeat(1, 1, 4)).sigmoid()
        tgt_undetach = torch.gather(om, 1, topk.unsqueeze(-1).repeat(1, 1, d))
        B = src.shape[0]
        tgt = sd["transformer.tgt_embed.weight"][None].repeat(B, 1, 1)         # embed_init_tgt=True
        refpoint = refpoint_undetach

        attn_mask = None
        dn_meta = None
        if targets is not None and training:
            q_label, q_bbox, attn_mask, dn_meta = prepare_for_cdn_nonoise(targets, sd, cfg)
            refpoint = torch.cat([q_bbox, refpoint], 1)
            tgt = torch.cat([q_label, tgt], 1)

        # ---- decoder (deformable_transformer.py:652-766), batch-first restatement
        reference_points = refpoint.sigmoid()
        refs = [reference_points]
        hs = []
        output = tgt
        vr4 = torch.cat([valid_ratios, valid_ratios], -1)
        for i in range(cfg.dec_layers):
            ref_in = reference_points[:, :, None] * vr4[:, None]
            sine = gen_sineembed_for_position(ref_in[:, :, 0, :])
            query_pos = _mlp(sd, "transformer.decoder.ref_point_head", sine, 2)
            lst = {} if i == 0 else None
            output = decoder_layer(sd, "transformer.decoder.layers.%d" % i, output, query_pos, ref_in, memory,
                                   shapes, lsi, pad, cfg, attn_mask, lst)
            if lst is not None:
                st["dec0"] = lst
            delta = _mlp(sd, "transformer.decoder.bbox_embed.%d" % i, output, 3)
            reference_points = (delta + inverse_sigmoid(reference_points)).sigmoid()
            refs.append(reference_points)
            hs.append(_ln(sd, "transformer.decoder.norm", output))
        st["hs"] = hs
        st["refs"] = refs

        # ---- heads (dino.py:339-354)
        coords = torch.stack([(_mlp(sd, "bbox_embed.%d" % i, hs[i], 3) + inverse_sigmoid(refs[i])).sigmoid()
                              for i in range(cfg.dec_layers)])
        classes = torch.stack([_linear(sd, "class_embed.%d" % i, hs[i]) for i in range(cfg.dec_layers)])
        # NB quirk Q3: dn_post_process is skipped when model.dn_number == 0 (dino.py:355) -> DN queries stay in the output
        res = {"pred_logits": classes[-1], "pred_boxes": coords[-1],
               "aux_outputs": [{"pred_logits": a, "pred_boxes": b} for a, b in zip(classes[:-1], coords[:-1])]}
        interm_class = _linear(sd, "transformer.enc_out_class_embed", tgt_undetach)
        res["interm_outputs"] = {"pred_logits": interm_class, "pred_boxes": refpoint_undetach.sigmoid()}
        res["interm_outputs_for_matching_pre"] = {"pred_logits": interm_class, "pred_boxes": init_box_proposal}
        res["dn_meta"] = dn_meta
        return res


# ----------------------------------------------------------------------------------------------- decode tail
def ctc_view(pred_logits, pred_boxes, eps=0.003):
    """reference models/dino/dino.py:472-502: sort by cx, sigmoid, synthesise the blank column.
    Returns new_pred_logits (B,Q,C+1) and the sort permutation."""
    _, idx = torch.sort(pred_boxes[:, :, 0])
    p = torch.gather(pred_logits, 1, idx.unsqueeze(-1).expand(-1, -1, pred_logits.shape[-1])).sigmoid()
    s = p.sum(-1)
    low = s < 1 - eps
    out = torch.zeros(p.shape[0], p.shape[1], p.shape[2] + 1, dtype=p.dtype)
    out[..., 1:] = torch.where(low[..., None], p, (1 - eps) * p / s[..., None])
    out[..., 0] = torch.where(low, 1 - s, torch.full_like(s, eps))
    return out, idx


def ctc_decode(new_pred_logits):
    """reference engine.py:512-530 (convert_output_to_pred, duplicate=False): argmax, drop blanks, -1."""
    am = new_pred_logits.argmax(-1)
    return [(row[row != 0] - 1).tolist() for row in am]


def loss_ctc(pred_logits, pred_boxes, targets, eps=0.003):
    """reference models/dino/dino.py:457-551: hard-blank interleave + nn.CTCLoss(blank=0, zero_infinity, mean)."""
    new, _ = ctc_view(pred_logits, pred_boxes, eps)
    B, Q, C1 = new.shape
    blank = torch.full_like(new, 1e-5)
    blank[..., 0] = 1
    padded = torch.zeros(B, 2 * Q, C1, dtype=new.dtype)
    padded[:, ::2] = new
    padded[:, 1::2] = blank
    lens = torch.tensor([len(t["labels"]) for t in targets], dtype=torch.int64)
    tt = torch.zeros(B, int(lens.max()))
    for i, t in enumerate(targets):
        tt[i, : len(t["labels"])] = t["labels"] + 1
    loss = F.ctc_loss(torch.log(padded.permute(1, 0, 2)), tt, torch.full((B,), 2 * Q, dtype=torch.int64), lens,
                      blank=0, reduction="mean", zero_infinity=True)
    return loss, new


def postprocess(pred_logits, pred_boxes, target_sizes, num_select, not_to_xyxy=False):
    """reference models/dino/dino.py:994-1046 without NMS (nms_iou_threshold <= 0)."""
    B, Q, C = pred_logits.shape
    prob = pred_logits.sigmoid()
    scores, idx = torch.topk(prob.view(B, -1), num_select, dim=1)
    qi = idx // C
    labels = idx % C
    if not_to_xyxy:
        boxes = pred_boxes
    else:
        cx, cy, w, h = pred_boxes.unbind(-1)
        boxes = torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)
    boxes = torch.gather(boxes, 1, qi.unsqueeze(-1).repeat(1, 1, 4))
    img_h, img_w = target_sizes.unbind(1)
    boxes = boxes * torch.stack([img_w, img_h, img_w, img_h], 1)[:, None, :]
    return scores, labels, boxes

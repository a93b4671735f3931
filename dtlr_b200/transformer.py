"""Deformable transformer: the parameter-holding modules with the reference's state-dict layout and call signature
(reference models/dino/deformable_transformer.py:25-1069), restricted to what the shipped DTLR configs use:
deformable encoder + deformable decoder, two_stage_type in {'standard','no'}, decoder_sa_type 'sa',
module_seq ['sa','ca','ffn'], post-norm, rm_dec_query_scale, embed_init_tgt.

`forward` here is the autograd-capable module path (training, fine-tuning, and the reference-shaped eval path):
dense contractions go through torch (cuBLAS), the deformable-attention core through the C ABI.  The fused inference
path lives in dtlr_b200/engine.py and reads the same parameters.
"""
import copy
import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn, Tensor

from .misc import inverse_sigmoid
from .ms_deform_attn import MSDeformAttn


class MLP(nn.Module):
    """reference models/dino/utils.py:110-122"""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = F.relu(layer(x)) if i < self.num_layers - 1 else layer(x)
        return x


def _activation(name):
    if name == "relu":
        return F.relu
    if name == "gelu":
        return F.gelu
    raise RuntimeError("activation should be relu/gelu, not %s." % name)


def gen_sineembed_for_position(pos_tensor):
    """reference models/dino/utils.py:141-167: order (y, x, w, h), T=10000, scale 2*pi."""
    scale = 2 * math.pi
    dim_t = torch.arange(128, dtype=torch.float32, device=pos_tensor.device)
    dim_t = 10000 ** (2 * (dim_t // 2) / 128)

    def emb(v):
        e = (v * scale)[..., None] / dim_t
        return torch.stack((e[..., 0::2].sin(), e[..., 1::2].cos()), dim=-1).flatten(-2)

    parts = [emb(pos_tensor[..., 1]), emb(pos_tensor[..., 0])]
    if pos_tensor.size(-1) == 4:
        parts += [emb(pos_tensor[..., 2]), emb(pos_tensor[..., 3])]
    elif pos_tensor.size(-1) != 2:
        raise ValueError("Unknown pos_tensor shape(-1):{}".format(pos_tensor.size(-1)))
    return torch.cat(parts, dim=-1)


def gen_encoder_output_proposals(memory: Tensor, memory_padding_mask: Tensor, spatial_shapes_list, default_hw=0.05):
    """reference models/dino/utils.py:15-64; spatial_shapes_list is a host list [(H,W),...] (no device sync)."""
    N_, S_, C_ = memory.shape
    proposals = []
    _cur = 0
    for lvl, (H_, W_) in enumerate(spatial_shapes_list):
        m = memory_padding_mask[:, _cur:(_cur + H_ * W_)].view(N_, H_, W_, 1)
        valid_H = torch.sum(~m[:, :, 0, 0], 1)
        valid_W = torch.sum(~m[:, 0, :, 0], 1)
        gy, gx = torch.meshgrid(torch.linspace(0, H_ - 1, H_, dtype=torch.float32, device=memory.device),
                                torch.linspace(0, W_ - 1, W_, dtype=torch.float32, device=memory.device), indexing="ij")
        grid = torch.cat([gx.unsqueeze(-1), gy.unsqueeze(-1)], -1)
        scale = torch.cat([valid_W.unsqueeze(-1), valid_H.unsqueeze(-1)], 1).view(N_, 1, 1, 2)
        grid = (grid.unsqueeze(0).expand(N_, -1, -1, -1) + 0.5) / scale
        wh = torch.ones_like(grid) * default_hw * (2.0 ** lvl)
        proposals.append(torch.cat((grid, wh), -1).view(N_, -1, 4))
        _cur += H_ * W_
    prop = torch.cat(proposals, 1)
    valid = ((prop > 0.01) & (prop < 0.99)).all(-1, keepdim=True)
    prop = torch.log(prop / (1 - prop))
    prop = prop.masked_fill(memory_padding_mask.unsqueeze(-1), float("inf"))
    prop = prop.masked_fill(~valid, float("inf"))
    out_mem = memory.masked_fill(memory_padding_mask.unsqueeze(-1), float(0)).masked_fill(~valid, float(0))
    return out_mem, prop


class DeformableTransformerEncoderLayer(nn.Module):
    """reference deformable_transformer.py:768-823"""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _activation(activation)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, key_padding_mask=None):
        src2 = self.self_attn(src if pos is None else src + pos, reference_points, src, spatial_shapes,
                              level_start_index, key_padding_mask)
        src = self.norm1(src + self.dropout1(src2))
        src2 = self.linear2(self.dropout2(self.activation(self.linear1(src))))
        return self.norm2(src + self.dropout3(src2))


class TransformerEncoder(nn.Module):
    """reference deformable_transformer.py:437-580 (deformable, two_stage_type 'no'/'standard')"""

    def __init__(self, encoder_layer, num_layers, d_model=256):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)]) if num_layers > 0 else []
        self.num_layers = num_layers
        self.d_model = d_model
        self.norm = None

    @staticmethod
    def get_reference_points(spatial_shapes_list, valid_ratios, device):
        """reference :479-492 with the level loop over a host list (the reference iterates a CUDA tensor -> syncs)."""
        pts = []
        for lvl, (H_, W_) in enumerate(spatial_shapes_list):
            ref_y, ref_x = torch.meshgrid(torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32, device=device),
                                          torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32, device=device), indexing="ij")
            ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H_)
            ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W_)
            pts.append(torch.stack((ref_x, ref_y), -1))
        ref = torch.cat(pts, 1)
        return ref[:, :, None] * valid_ratios[:, None]

    def forward(self, src, pos, spatial_shapes, level_start_index, valid_ratios, key_padding_mask,
                spatial_shapes_list=None):
        if spatial_shapes_list is None:
            spatial_shapes_list = [tuple(int(v) for v in r) for r in spatial_shapes.tolist()]
        output = src
        if self.num_layers > 0:
            reference_points = self.get_reference_points(spatial_shapes_list, valid_ratios, device=src.device)
        for layer in self.layers:
            output = layer(src=output, pos=pos, reference_points=reference_points, spatial_shapes=spatial_shapes,
                           level_start_index=level_start_index, key_padding_mask=key_padding_mask)
        return output, None, None


class DeformableTransformerDecoderLayer(nn.Module):
    """reference deformable_transformer.py:825-997, module_seq ['sa','ca','ffn'], decoder_sa_type 'sa'"""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4,
                 module_seq=("sa", "ca", "ffn")):
        super().__init__()
        self.module_seq = list(module_seq)
        assert sorted(self.module_seq) == ["ca", "ffn", "sa"]
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _activation(activation)
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)
        self.label_embedding = None

    def forward_ffn(self, tgt):
        tgt2 = self.linear2(self.dropout3(self.activation(self.linear1(tgt))))
        return self.norm3(tgt + self.dropout4(tgt2))

    def forward_sa(self, tgt, query_pos, self_attn_mask):
        q = k = tgt if query_pos is None else tgt + query_pos
        tgt2 = self.self_attn(q, k, tgt, attn_mask=self_attn_mask)[0]
        return self.norm2(tgt + self.dropout2(tgt2))

    def forward_ca(self, tgt, query_pos, ref_pts, memory, shapes, lsi, mem_mask):
        q = tgt if query_pos is None else tgt + query_pos
        tgt2 = self.cross_attn(q.transpose(0, 1), ref_pts.transpose(0, 1).contiguous(), memory.transpose(0, 1),
                               shapes, lsi, mem_mask).transpose(0, 1)
        return self.norm1(tgt + self.dropout1(tgt2))

    def forward(self, tgt, tgt_query_pos=None, tgt_query_sine_embed=None, tgt_key_padding_mask=None,
                tgt_reference_points=None, memory=None, memory_key_padding_mask=None, memory_level_start_index=None,
                memory_spatial_shapes=None, memory_pos=None, self_attn_mask=None, cross_attn_mask=None):
        for name in self.module_seq:
            if name == "ffn":
                tgt = self.forward_ffn(tgt)
            elif name == "ca":
                tgt = self.forward_ca(tgt, tgt_query_pos, tgt_reference_points, memory, memory_spatial_shapes,
                                      memory_level_start_index, memory_key_padding_mask)
            elif name == "sa":
                tgt = self.forward_sa(tgt, tgt_query_pos, self_attn_mask)
            else:
                raise ValueError("unknown funcname {}".format(name))
        return tgt


class TransformerDecoder(nn.Module):
    """reference deformable_transformer.py:582-766 (deformable decoder, query_dim 4, iterative box refinement)"""

    def __init__(self, decoder_layer, num_layers, norm=None, d_model=256, query_dim=4, num_feature_levels=1):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(decoder_layer) for _ in range(num_layers)]) if num_layers > 0 else []
        self.num_layers = num_layers
        self.norm = norm
        self.query_dim = query_dim
        self.num_feature_levels = num_feature_levels
        self.ref_point_head = MLP(query_dim // 2 * d_model, d_model, d_model, 2)
        self.query_pos_sine_scale = None
        self.query_scale = None
        self.ref_anchor_head = None
        self.bbox_embed = None
        self.class_embed = None
        self.d_model = d_model
        self.dec_layer_number = None
        self.rm_detach = None

    def forward(self, tgt, memory, tgt_mask: Optional[Tensor] = None, memory_key_padding_mask: Optional[Tensor] = None,
                pos: Optional[Tensor] = None, refpoints_unsigmoid: Optional[Tensor] = None,
                level_start_index: Optional[Tensor] = None, spatial_shapes: Optional[Tensor] = None,
                valid_ratios: Optional[Tensor] = None):
        output = tgt
        intermediate = []
        reference_points = refpoints_unsigmoid.sigmoid()
        ref_points = [reference_points]
        for layer_id, layer in enumerate(self.layers):
            if reference_points.shape[-1] == 4:
                ref_in = reference_points[:, :, None] * torch.cat([valid_ratios, valid_ratios], -1)[None, :]
            else:
                ref_in = reference_points[:, :, None] * valid_ratios[None, :]
            query_sine_embed = gen_sineembed_for_position(ref_in[:, :, 0, :])
            query_pos = self.ref_point_head(query_sine_embed)
            output = layer(tgt=output, tgt_query_pos=query_pos, tgt_query_sine_embed=query_sine_embed,
                           tgt_reference_points=ref_in, memory=memory, memory_key_padding_mask=memory_key_padding_mask,
                           memory_level_start_index=level_start_index, memory_spatial_shapes=spatial_shapes,
                           memory_pos=pos, self_attn_mask=tgt_mask)
            if self.bbox_embed is not None:
                new_ref = (self.bbox_embed[layer_id](output) + inverse_sigmoid(reference_points)).sigmoid()
                reference_points = new_ref.detach()
                ref_points.append(new_ref)
            intermediate.append(self.norm(output))
        return [[o.transpose(0, 1) for o in intermediate], [r.transpose(0, 1) for r in ref_points]]


class DeformableTransformer(nn.Module):
    """reference deformable_transformer.py:25-435"""

    def __init__(self, d_model=256, nhead=8, num_queries=300, num_encoder_layers=6, num_decoder_layers=6,
                 dim_feedforward=2048, dropout=0.0, activation="relu", num_feature_levels=4, enc_n_points=4,
                 dec_n_points=4, two_stage_type="standard", module_seq=("sa", "ca", "ffn"), embed_init_tgt=True,
                 two_stage_default_hw=0.05, query_dim=4):
        super().__init__()
        assert query_dim == 4
        assert two_stage_type in ["no", "standard"], "unknown param {} of two_stage_type".format(two_stage_type)
        self.num_feature_levels = num_feature_levels
        self.num_encoder_layers = num_encoder_layers
        self.num_decoder_layers = num_decoder_layers
        self.num_queries = num_queries
        self.d_model = d_model
        self.nhead = nhead
        self.dec_layers = num_decoder_layers
        self.two_stage_type = two_stage_type
        self.two_stage_default_hw = two_stage_default_hw
        self.two_stage_keep_all_tokens = False
        self.embed_init_tgt = embed_init_tgt
        enc_layer = DeformableTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation, num_feature_levels,
                                                      nhead, enc_n_points)
        self.encoder = TransformerEncoder(enc_layer, num_encoder_layers, d_model=d_model)
        dec_layer = DeformableTransformerDecoderLayer(d_model, dim_feedforward, dropout, activation, num_feature_levels,
                                                      nhead, dec_n_points, module_seq)
        self.decoder = TransformerDecoder(dec_layer, num_decoder_layers, nn.LayerNorm(d_model), d_model=d_model,
                                          query_dim=query_dim, num_feature_levels=num_feature_levels)
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model)) if num_feature_levels > 1 else None
        if (two_stage_type != "no" and embed_init_tgt) or two_stage_type == "no":
            self.tgt_embed = nn.Embedding(num_queries, d_model)
            nn.init.normal_(self.tgt_embed.weight.data)
        else:
            self.tgt_embed = None
        if two_stage_type == "standard":
            self.enc_output = nn.Linear(d_model, d_model)
            self.enc_output_norm = nn.LayerNorm(d_model)
            self.two_stage_wh_embedding = None
        if two_stage_type == "no":
            self.refpoint_embed = nn.Embedding(num_queries, 4)
        self.enc_out_class_embed = None
        self.enc_out_bbox_embed = None
        self._reset_parameters()
        self.debug_force_topk = None   # test hook: (B,Q) int64 indices replacing the two-stage top-k (parity protocol)
        self.debug_stages = None       # test hook: dict filled with intermediates

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        if self.level_embed is not None:
            nn.init.normal_(self.level_embed)

    @staticmethod
    def get_valid_ratio(mask):
        _, H, W = mask.shape
        valid_H = torch.sum(~mask[:, :, 0], 1)
        valid_W = torch.sum(~mask[:, 0, :], 1)
        return torch.stack([valid_W.float() / W, valid_H.float() / H], -1)

    def forward(self, srcs, masks, refpoint_embed, pos_embeds, tgt, attn_mask=None):
        src_flatten, mask_flatten, pos_flatten, shapes_list = [], [], [], []
        for lvl, (src, mask, pos_embed) in enumerate(zip(srcs, masks, pos_embeds)):
            bs, c, h, w = src.shape
            shapes_list.append((h, w))
            src_flatten.append(src.flatten(2).transpose(1, 2))
            mask_flatten.append(mask.flatten(1))
            pe = pos_embed.flatten(2).transpose(1, 2)
            if self.level_embed is not None:
                pe = pe + self.level_embed[lvl].view(1, 1, -1)
            pos_flatten.append(pe)
        src_flatten = torch.cat(src_flatten, 1)
        mask_flatten = torch.cat(mask_flatten, 1)
        pos_flatten = torch.cat(pos_flatten, 1)
        spatial_shapes = torch.as_tensor(shapes_list, dtype=torch.long, device=src_flatten.device)
        starts = [0]
        for h, w in shapes_list[:-1]:
            starts.append(starts[-1] + h * w)
        level_start_index = torch.as_tensor(starts, dtype=torch.long, device=src_flatten.device)
        valid_ratios = torch.stack([self.get_valid_ratio(m) for m in masks], 1)

        memory, _, _ = self.encoder(src_flatten, pos=pos_flatten, level_start_index=level_start_index,
                                    spatial_shapes=spatial_shapes, valid_ratios=valid_ratios,
                                    key_padding_mask=mask_flatten, spatial_shapes_list=shapes_list)
        st = self.debug_stages
        if st is not None:
            st["memory"] = memory

        if self.two_stage_type == "standard":
            output_memory, output_proposals = gen_encoder_output_proposals(memory, mask_flatten, shapes_list,
                                                                           self.two_stage_default_hw)
            output_memory = self.enc_output_norm(self.enc_output(output_memory))
            cls_unsel = self.enc_out_class_embed(output_memory)
            coord_unsel = self.enc_out_bbox_embed(output_memory) + output_proposals
            scores = cls_unsel.max(-1)[0]
            topk_proposals = torch.topk(scores, self.num_queries, dim=1)[1]
            if self.debug_force_topk is not None:
                topk_proposals = self.debug_force_topk.to(scores.device)
            if st is not None:
                st["topk_scores"], st["topk_idx"] = scores, topk_proposals
            refpoint_embed_undetach = torch.gather(coord_unsel, 1, topk_proposals.unsqueeze(-1).repeat(1, 1, 4))
            refpoint_embed_ = refpoint_embed_undetach.detach()
            init_box_proposal = torch.gather(output_proposals, 1, topk_proposals.unsqueeze(-1).repeat(1, 1, 4)).sigmoid()
            tgt_undetach = torch.gather(output_memory, 1, topk_proposals.unsqueeze(-1).repeat(1, 1, self.d_model))
            if self.embed_init_tgt:
                tgt_ = self.tgt_embed.weight[:, None, :].repeat(1, bs, 1).transpose(0, 1)
            else:
                tgt_ = tgt_undetach.detach()
            if refpoint_embed is not None:
                refpoint_embed = torch.cat([refpoint_embed, refpoint_embed_], dim=1)
                tgt = torch.cat([tgt, tgt_], dim=1)
            else:
                refpoint_embed, tgt = refpoint_embed_, tgt_
        else:
            tgt_ = self.tgt_embed.weight[:, None, :].repeat(1, bs, 1).transpose(0, 1)
            refpoint_embed_ = self.refpoint_embed.weight[:, None, :].repeat(1, bs, 1).transpose(0, 1)
            if refpoint_embed is not None:
                refpoint_embed = torch.cat([refpoint_embed, refpoint_embed_], dim=1)
                tgt = torch.cat([tgt, tgt_], dim=1)
            else:
                refpoint_embed, tgt = refpoint_embed_, tgt_
            init_box_proposal = refpoint_embed_.sigmoid()

        hs, references = self.decoder(tgt=tgt.transpose(0, 1), memory=memory.transpose(0, 1),
                                      memory_key_padding_mask=mask_flatten, pos=pos_flatten.transpose(0, 1),
                                      refpoints_unsigmoid=refpoint_embed.transpose(0, 1),
                                      level_start_index=level_start_index, spatial_shapes=spatial_shapes,
                                      valid_ratios=valid_ratios, tgt_mask=attn_mask)
        if self.two_stage_type == "standard":
            hs_enc = tgt_undetach.unsqueeze(0)
            ref_enc = refpoint_embed_undetach.sigmoid().unsqueeze(0)
        else:
            hs_enc = ref_enc = None
        return hs, references, hs_enc, ref_enc, init_box_proposal


def build_deformable_transformer(args):
    """reference deformable_transformer.py:1007-1069"""
    unsupported = dict(decoder_layer_noise=False, use_deformable_box_attn=False, add_channel_attention=False,
                       two_stage_pat_embed=0, two_stage_add_query_num=0, two_stage_learn_wh=False,
                       two_stage_keep_all_tokens=False, dec_layer_number=None, num_patterns=0, pre_norm=False,
                       decoder_sa_type="sa")
    for k, v in unsupported.items():
        if getattr(args, k, v) != v:
            raise NotImplementedError("dtlr_b200: config option %s=%r is not used by any shipped DTLR config" % (k, getattr(args, k)))
    return DeformableTransformer(
        d_model=args.hidden_dim, dropout=args.dropout, nhead=args.nheads, num_queries=args.num_queries,
        dim_feedforward=args.dim_feedforward, num_encoder_layers=args.enc_layers, num_decoder_layers=args.dec_layers,
        query_dim=args.query_dim, activation=args.transformer_activation, num_feature_levels=args.num_feature_levels,
        enc_n_points=args.enc_n_points, dec_n_points=args.dec_n_points, two_stage_type=args.two_stage_type,
        module_seq=args.decoder_module_seq, embed_init_tgt=args.embed_init_tgt,
        two_stage_default_hw=getattr(args, "two_stage_default_hw", 0.05))

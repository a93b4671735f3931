"""Native fine-tune step: BASELINE config 5 = reference engine.py:172-274 `train_one_epoch_CTC`
(forward(samples, targets) -> SetCriterion.loss_CTC -> loss.backward() -> clip_grad_norm_ -> AdamW.step, DDP all-reduce inside
backward, finetuning.py:211-231) with the transformer -- encoder, two-stage selection, decoder, heads, loss, optimizer -- running on
libdtlr_b200 kernels in BOTH directions.  No autograd graph is built for that part: the backward is an explicit launch sequence
(below), mirrored statement by statement on the forward.

What runs where
  * front (ResNet-50 layer2-4 + input_proj + GroupNorm; reference backbone.py:109-128, dino.py:118-135, 304-318): torch modules under
    autograd (cuDNN); the native part hands back d loss / d src_flatten.  [native conv dgrad / wgrad: DESIGN.md "training step"]
  * encoder x6, decoder x6, class head (deformable_transformer.py:437-766, dino.py:339-346): tcgen05 GEMMs forward, dgrad = the same
    GEMM kernel against the packed W^T, wgrad = dtlr_wgrad (MN-major tcgen05, split over rows, fp32 reductions into the gradient
    arena), LayerNorm / ReLU / MSDeformAttn prologue backward kernels, the MSDA core forward / backward kernels.
  * loss_CTC forward + backward: the fused lattice kernels (dtlr_ctc_loss).
  * clip_grad_norm_ + AdamW: two kernels over flat fp32 arenas (parameters, gradients, both moments); one more launch re-creates the
    16-bit W / W^T operand copies of every Linear.
  * N > 1: one process per GPU; the gradient arena is ordered decoder | encoder | front so each segment is all-reduced (NCCL, async)
    as soon as its backward has been issued, overlapping the rest of the backward.  1 / world is folded into the loss gradient.

Which gradients exist.  In the CTC step only pred_logits of the LAST decoder layer enters the loss (pred_boxes steers the cx sort);
reference points are detached between layers (deformable_transformer.py:737) and tgt is the learned table (embed_init_tgt), so
enc_output / enc_output_norm / enc_out_* / bbox_embed receive no gradient: the reference leaves their .grad at None (DDP
find_unused_parameters, finetuning.py:211-215) and AdamW skips them -- they are not part of the arena here either.

Precision: operands of every contraction in `dtype` (bf16: tcgen05; fp32: the exact SIMT kernels = parity mode), accumulation,
LayerNorm statistics, the MSDA core, the loss, the residual GRADIENT stream, the gradient arena and the optimizer in fp32.
"""
import math

import torch
import torch.nn.functional as F

from .dino import prepare_for_cdn
from .misc import NestedTensor, nested_tensor_from_tensor_list

# (named_parameters() reports the shared heads under their first alias: transformer.decoder.{bbox,class}_embed.0.*)
DEAD_PREFIXES = ("transformer.enc_output", "transformer.enc_out_", "bbox_embed.", "transformer.decoder.bbox_embed.")
HEAD_PREFIXES = ("class_embed.", "transformer.decoder.class_embed.")
F32 = torch.float32


class Lin:
    """operand copies (w16 [N, K], wT16 [K, ldn]) and master / gradient views of one -- possibly fused -- nn.Linear"""

    def __init__(self, name, parts, dtype, device, need_T=True):
        # parts: list of (weight_param, row0, nrows, bias_param | None)
        self.name, self.parts = name, parts
        self.N = sum(p[2] for p in parts)
        self.K = parts[0][0].shape[1]
        pad = 8 if dtype != F32 else 4
        self.ldn = (self.N + pad - 1) // pad * pad
        self.w16 = torch.zeros((self.N, self.K), dtype=dtype, device=device)
        self.wT16 = torch.zeros((self.K, self.ldn), dtype=dtype, device=device) if need_T else None
        self._bias = None
        self.fused_bias = None
        self.master_w = parts[0][0] if len(parts) == 1 and parts[0][1] == 0 and parts[0][2] == parts[0][0].shape[0] else None
        self.train = any(p[0].requires_grad for p in parts)

    @property
    def bias(self):
        """fp32 bias of the fused rows (a view of the master parameter when there is one part, else a packed copy)"""
        if len(self.parts) == 1:
            w, r0, n, b = self.parts[0]
            return None if b is None else b.detach()[r0:r0 + n]
        return self.fused_bias

    def refresh_bias(self):
        if len(self.parts) > 1:
            self.fused_bias = torch.cat([b.detach()[r0:r0 + n] for _, r0, n, b in self.parts]).contiguous()


class TrainEngine:
    def __init__(self, model, param_groups=None, lr=1e-4, lr_backbone=1e-5, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8,
                 max_norm=0.1, dtype=torch.bfloat16, K=None, process_group=None, world_size=1, native_front=None):
        if native_front is None:
            native_front = K is None            # the torch stand-in of tests/ keeps the front under autograd
        if K is None:
            from . import train_ops as K_
            K = K_
        self.K = K
        self.model = model
        self.T = dtype
        self.betas, self.eps, self.max_norm = betas, eps, max_norm
        self.world, self.pg = world_size, process_group
        tr = model.transformer
        if not (tr.two_stage_type == "standard" and tr.embed_init_tgt):
            raise NotImplementedError("TrainEngine: two_stage_type 'standard' with embed_init_tgt (every shipped DTLR config)")
        self.device = next(model.parameters()).device
        if param_groups is None:        # reference util/get_param_dicts.py:23-31 ('default')
            named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
            param_groups = [{"params": [p for n, p in named if "backbone" not in n], "lr": lr, "weight_decay": weight_decay},
                            {"params": [p for n, p in named if "backbone" in n], "lr": lr_backbone, "weight_decay": weight_decay}]
        self._build_arena(param_groups, lr, weight_decay)
        self._build_linears()
        self.front = None
        if native_front:
            from .train_front import NativeFront
            self.front = NativeFront(self)          # (packs its own operand copies)
        self.repack(front=False)
        self.last = {}
        self._masks = {}

    # ------------------------------------------------------------------------------------------------ arenas
    def _segment_of(self, name):
        # segment 2 = everything whose gradient is produced by the closing torch.autograd.backward call: the front, and the two
        # embedding tables behind the decoder's input (tgt_embed, label_enc) -- NOT final when the decoder backward has been issued
        if name.startswith(("backbone.", "input_proj.", "transformer.tgt_embed.", "label_enc.")):
            return 2
        if name.startswith("transformer.encoder.") or name == "transformer.level_embed":
            return 1
        return 0

    def _build_arena(self, param_groups, lr, weight_decay):
        """flat fp32 arenas: parameters (the nn.Parameters are re-pointed at views), gradients (p.grad = view), exp_avg, exp_avg_sq.
        Order: optimizer group, then backward completion order (decoder + heads | encoder | front)."""
        names = {id(p): n for n, p in self.model.named_parameters()}
        entries = []
        for gi, g in enumerate(param_groups):
            for p in g["params"]:
                n = names.get(id(p))
                if n is None or not p.requires_grad or n.startswith(DEAD_PREFIXES):
                    continue
                entries.append((gi, self._segment_of(n), len(entries), n, p))
        entries.sort(key=lambda e: (e[0], e[1], e[2]))
        off = 0
        layout = []
        for gi, seg, _, n, p in entries:
            layout.append((n, p, gi, seg, off, p.numel()))
            off += (p.numel() + 3) // 4 * 4          # 16-byte aligned views (vector reductions of dtlr_wgrad)
        total = off
        dev = self.device
        self.flat_p = torch.zeros(total, dtype=F32, device=dev)
        self.flat_g = torch.zeros(total, dtype=F32, device=dev)
        self.flat_m = torch.zeros(total, dtype=F32, device=dev)
        self.flat_v = torch.zeros(total, dtype=F32, device=dev)
        self.state = torch.zeros(4, dtype=F32, device=dev)
        self.layout = layout
        self.grad_of = {}
        with torch.no_grad():
            for n, p, gi, seg, o, k in layout:
                view = self.flat_p[o:o + k].view(p.shape)
                view.copy_(p.detach())
                p.data = view
                p.grad = self.flat_g[o:o + k].view(p.shape)
                self.grad_of[id(p)] = p.grad
        # optimizer ranges (one per group) and all-reduce chunks (one per (group, segment) run)
        # `param_groups` mirrors torch.optim.Optimizer.param_groups (lr / weight_decay are read at every step, so a scheduler -- or
        # the reference's per-epoch lr drop, finetuning.py:616-622 -- only has to write g["lr"])
        self.group_ranges, self.chunks, self.param_groups = [], [], []
        for gi, g in enumerate(param_groups):
            idx = [e for e in layout if e[2] == gi]
            if not idx:
                continue
            a, b = idx[0][4], idx[-1][4] + (idx[-1][5] + 3) // 4 * 4
            pg = {"lr": float(g.get("lr", lr)), "weight_decay": float(g.get("weight_decay", weight_decay)), "params": [e[1] for e in idx]}
            pg["initial_lr"] = pg["lr"]
            self.param_groups.append(pg)
            self.group_ranges.append((a, b, pg))
        run = None
        for n, p, gi, seg, o, k in layout:
            end = o + (k + 3) // 4 * 4
            if run is not None and run[0] == (gi, seg):
                run[2] = end
            else:
                if run is not None:
                    self.chunks.append((run[0][1], run[1], run[2]))
                run = [(gi, seg), o, end]
        if run is not None:
            self.chunks.append((run[0][1], run[1], run[2]))
        self.n_live = sum(e[5] for e in layout)
        if getattr(self.model, "_engine", None) is not None:
            self.model.invalidate_engine()

    def grad(self, p):
        return self.grad_of.get(id(p))

    # ------------------------------------------------------------------------------------------------ linears
    def _lin(self, name, parts, need_T=True):
        l = Lin(name, parts, self.T, self.device, need_T)
        self.all_lins.append(l)
        return l

    def _build_linears(self):
        m, tr = self.model, self.model.transformer
        self.all_lins = []
        d = tr.d_model

        def one(name, mod, need_T=True):
            return self._lin(name, [(mod.weight, 0, mod.weight.shape[0], mod.bias)], need_T)

        def msda(name, a):
            return {"val": one(name + ".val", a.value_proj),
                    "oa": self._lin(name + ".oa", [(a.sampling_offsets.weight, 0, a.sampling_offsets.weight.shape[0], a.sampling_offsets.bias),
                                                   (a.attention_weights.weight, 0, a.attention_weights.weight.shape[0], a.attention_weights.bias)]),
                    "out": one(name + ".out", a.output_proj), "M": a.n_heads, "L": a.n_levels, "P": a.n_points}

        self.enc = [{"attn": msda("enc%d" % i, l.self_attn), "ln1": l.norm1, "l1": one("enc%d.l1" % i, l.linear1),
                     "l2": one("enc%d.l2" % i, l.linear2), "ln2": l.norm2} for i, l in enumerate(tr.encoder.layers)]
        self.dec = []
        for i, l in enumerate(tr.decoder.layers):
            sa = l.self_attn
            self.dec.append({"ca": msda("dec%d" % i, l.cross_attn), "ln1": l.norm1,
                             "qk": self._lin("dec%d.qk" % i, [(sa.in_proj_weight, 0, 2 * d, sa.in_proj_bias)]),
                             "v": self._lin("dec%d.v" % i, [(sa.in_proj_weight, 2 * d, d, sa.in_proj_bias)]),
                             "o": one("dec%d.o" % i, sa.out_proj), "heads": sa.num_heads, "ln2": l.norm2,
                             "l1": one("dec%d.l1" % i, l.linear1), "l2": one("dec%d.l2" % i, l.linear2), "ln3": l.norm3})
        self.rph = [one("rph%d" % i, l) for i, l in enumerate(tr.decoder.ref_point_head.layers)]
        self.lin = {"enc_output": one("enc_output", tr.enc_output, False), "enc_cls": one("enc_cls", tr.enc_out_class_embed, False),
                    "enc_bbox": [one("enc_bbox%d" % i, l, False) for i, l in enumerate(tr.enc_out_bbox_embed.layers)]}
        self.bbox = [[one("bbox%d.%d" % (i, j), l, False) for j, l in enumerate(be.layers)] for i, be in enumerate(m.bbox_embed)]
        self.cls = one("cls", m.class_embed[-1])
        # device table of dtlr_pack_weights: one entry per part
        rows = []
        tiles = [0]
        for l in self.all_lins:
            r_acc = 0
            for w, r0, n, b in l.parts:
                src = w.detach()[r0:r0 + n]
                esz = l.w16.element_size()
                rows.append([src.data_ptr(), n, l.K, w.shape[1], l.w16.data_ptr() + r_acc * l.K * esz, l.K,
                             (l.wT16.data_ptr() + r_acc * esz) if l.wT16 is not None else 0, l.ldn])
                tiles.append(tiles[-1] + ((n + 31) // 32) * ((l.K + 31) // 32))
                r_acc += n
        self.pack_table = torch.tensor(rows, dtype=torch.int64).to(self.device)
        self.pack_tiles = torch.tensor(tiles, dtype=torch.int32).to(self.device)
        self.pack_n, self.pack_total = len(rows), tiles[-1]

    def repack(self, front=True):
        """16-bit (or fp32) operand copies of every Linear / convolution from the fp32 masters: after construction and after every
        optimizer step"""
        if front and self.front is not None:
            self.front.repack()
        if hasattr(self.K, "repack_lins"):          # (the torch stand-in of tests/ has no raw-pointer table walker)
            self.K.repack_lins(self.all_lins, self.T)
        else:
            self.K.pack_weights(self.pack_table, self.pack_tiles, self.pack_n, self.pack_total, self.T)
        for l in self.all_lins:
            l.refresh_bias()

    @staticmethod
    def ln_params(norm):
        return norm.weight.detach(), norm.bias.detach()

    # ------------------------------------------------------------------------------------------------ small helpers
    def _fwd(self, x, lin, relu=0, residual=None, out_dtype=None):
        return self.K.gemm(x, lin.w16, lin.bias, residual=residual, relu=relu, out_dtype=out_dtype)

    def _bwd(self, lin, dy, x, need_dx=True, residual=None, out_dtype=None, relu_mask=None):
        """weight / bias gradients of `lin` into the arena; returns dX = dY . W (+ residual) when need_dx.  relu_mask = the saved
        output h of a ReLU that produced x: the dgrad epilogue masks its result (h > 0 ? dX : 0) -- no separate ReLU-backward pass"""
        K = self.K
        c0 = 0
        for w, r0, n, b in lin.parts:
            if w.requires_grad and self.grad(w) is not None:
                K.wgrad(dy[:, c0:c0 + n], x, self.grad(w)[r0:r0 + n])
            if b is not None and b.requires_grad and self.grad(b) is not None:
                K.colsum(dy[:, c0:c0 + n], self.grad(b)[r0:r0 + n])
            c0 += n
        if not need_dx:
            return None
        if relu_mask is not None:
            assert residual is None
            return K.gemm(dy, lin.wT16[:, :lin.N], None, residual=relu_mask, relu=3, out_dtype=relu_mask.dtype)
        return K.gemm(dy, lin.wT16[:, :lin.N], None, residual=residual, out_dtype=out_dtype)

    def _ln_bwd(self, norm, z, dy, dy2=None, want32=True, want16=True):
        g = self.grad(norm.weight) if norm.weight.requires_grad else None
        b = self.grad(norm.bias) if norm.bias.requires_grad else None
        return self.K.layernorm_bwd(z, dy, dy2, norm.weight.detach(), g, b, want32=want32, want16=want16, eps=norm.eps)

    # ------------------------------------------------------------------------------------------------ front (torch autograd)
    def _front(self, samples):
        """backbone + input_proj under autograd, position embeddings / masks without (reference dino.py:300-318,
        deformable_transformer.py:270-300).  Returns src_flatten (B,S,d) fp32 with grad_fn, pos (B,S,d) fp32 incl. level_embed
        (detached), mask_flatten, level shapes, masks."""
        m, tr = self.model, self.model.transformer
        with torch.enable_grad():
            features, poss = m.backbone(samples)
            srcs, masks = [], []
            for l, feat in enumerate(features):
                src, mask = feat.decompose()
                srcs.append(m.input_proj[l](src))
                masks.append(mask)
            poss = [p.detach() for p in poss]
            if m.num_feature_levels > len(srcs):
                n0 = len(srcs)
                for l in range(n0, m.num_feature_levels):
                    src = m.input_proj[l](features[-1].tensors) if l == n0 else m.input_proj[l](srcs[-1])
                    mask = F.interpolate(samples.mask[None].float(), size=src.shape[-2:]).to(torch.bool)[0]
                    with torch.no_grad():
                        poss.append(m.backbone[1](NestedTensor(src.detach(), mask)).to(src.dtype))
                    srcs.append(src)
                    masks.append(mask)
            level_hw = [tuple(s.shape[-2:]) for s in srcs]
            src_flatten = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
        with torch.no_grad():
            pos = torch.cat([p.flatten(2).transpose(1, 2) + tr.level_embed.detach()[l].view(1, 1, -1) for l, p in enumerate(poss)], 1)
            mask_flatten = torch.cat([mk.flatten(1) for mk in masks], 1)
        return src_flatten, pos, mask_flatten, level_hw, masks

    def _geometry(self, B, level_hw, masks, mask_flatten, nopad):
        from . import _lib as L
        tr = self.model.transformer
        S = sum(h * w for h, w in level_hw)
        starts = [0]
        for h, w in level_hw[:-1]:
            starts.append(starts[-1] + h * w)
        with torch.no_grad():
            vr = torch.stack([tr.get_valid_ratio(mk) for mk in masks], 1).float().contiguous()          # (B,L,2) = (w,h)
            vh = torch.stack([(~mk[:, :, 0]).sum(1) for mk in masks], 1)
            vw = torch.stack([(~mk[:, 0, :]).sum(1) for mk in masks], 1)
            valid_hw = torch.stack([vh, vw], -1).to(torch.int32).contiguous()
            pad_u8 = mask_flatten.contiguous().view(torch.uint8).reshape(-1)
        return {"B": B, "S": S, "nlev": len(level_hw), "level_hw": level_hw, "starts": starts,
                "shapes_host": L.i64_host([v for hw in level_hw for v in hw]), "lsi_host": L.i64_host(starts),
                "vr": vr, "valid_hw": valid_hw, "pad_u8": pad_u8, "pad_rows": None if nopad else pad_u8}

    # ------------------------------------------------------------------------------------------------ encoder
    def _enc_layer_fwd(self, lw, src, q, pos, geo, last):
        K, T = self.K, self.T
        a = lw["attn"]
        B, S = geo["B"], geo["S"]
        val = self._fwd(src, a["val"], out_dtype=F32)
        if geo["pad_rows"] is not None:
            K.zero_masked_rows_(val, geo["pad_rows"])
        oa = self._fwd(q, a["oa"], out_dtype=F32)
        loc, attn = K.msda_prep(oa, geo["ref_enc"], geo["vr"], geo, S, a["M"], a["P"])
        val4 = val.view(B, S, a["M"], -1)
        core = K.cast(K.msda_forward(val4, loc, attn, geo), T)
        z1 = self._fwd(core, a["out"], residual=src)
        s1 = K.layernorm(z1, *self.ln_params(lw["ln1"]))
        h = self._fwd(s1, lw["l1"], relu=1)
        z2 = self._fwd(h, lw["l2"], residual=s1)
        if last:
            out, qn = K.layernorm(z2, *self.ln_params(lw["ln2"])), None
        else:
            out, qn = K.layernorm(z2, *self.ln_params(lw["ln2"]), add2=pos)
        return out, qn, dict(src=src, q=q, val4=val4, loc=loc, attn=attn, core=core, z1=z1, s1=s1, h=h, z2=z2)

    def _enc_layer_bwd(self, lw, sv, d_y, d_qn, geo, need_dx, lvl_grad):
        """d_y: grad wrt the layer output (next layer's src), d_qn: grad wrt out + pos (next layer's query) or None.
        Returns (grad wrt this layer's src, grad wrt this layer's q)."""
        K, T = self.K, self.T
        a = lw["attn"]
        B, S = geo["B"], geo["S"]
        dz2_32, dz2 = self._ln_bwd(lw["ln2"], sv["z2"], d_y, d_qn)
        dh = self._bwd(lw["l2"], dz2, sv["h"], relu_mask=sv["h"])
        ds1 = self._bwd(lw["l1"], dh, sv["s1"], residual=dz2_32, out_dtype=F32)
        del dh
        dz1_32, dz1 = self._ln_bwd(lw["ln1"], sv["z1"], ds1)
        dcore = self._bwd(a["out"], dz1, sv["core"], out_dtype=F32)
        gv, gl, ga = K.msda_backward(sv["val4"], sv["loc"], sv["attn"], dcore.view(B, S, -1), geo)
        doa = K.msda_bwd_glue(gl, ga, sv["attn"], geo["ref_enc"], geo["vr"], geo, S, a["M"], a["P"], T)
        dq = self._bwd(a["oa"], doa, sv["q"], need_dx=need_dx or lvl_grad is not None, out_dtype=F32)
        if lvl_grad is not None:                     # level_embed: pos = sine + level_embed[l] enters every layer's query
            for l, (h_, w_) in enumerate(geo["level_hw"]):
                K.colsum(dq, lvl_grad[l], nseg=B, seg_rows=h_ * w_, seg_stride=S, row0=geo["starts"][l])
        gv2 = gv.view(B * S, -1)
        if geo["pad_rows"] is not None:
            K.zero_masked_rows_(gv2, geo["pad_rows"])
        dval = K.cast(gv2, T)
        dsrc = self._bwd(a["val"], dval, sv["src"], need_dx=need_dx, residual=dz1_32, out_dtype=F32)
        return dsrc, dq

    # ------------------------------------------------------------------------------------------------ decoder
    def _dec_layer_fwd(self, i, lw, tgt, ref, memory, mask_bool, geo, Qt):
        K, T = self.K, self.T
        B, S = geo["B"], geo["S"]
        a = lw["ca"]
        sine = K.sine_embed(ref, geo["vr"], B, Qt, geo["nlev"], T)
        r1 = self._fwd(sine, self.rph[0], relu=1)
        qp = self._fwd(r1, self.rph[1])
        qk_in = K.add(tgt, qp)
        qk = self._fwd(qk_in, lw["qk"])
        v = self._fwd(tgt, lw["v"])
        att, sactx = K.sa_forward(qk, v, mask_bool, B, Qt, lw["heads"])
        z2 = self._fwd(att, lw["o"], residual=tgt)
        t2, qca = K.layernorm(z2, *self.ln_params(lw["ln2"]), add2=qp)
        oa = self._fwd(qca, a["oa"], out_dtype=F32)
        loc, attn = K.msda_prep(oa, ref, geo["vr"], geo, Qt, a["M"], a["P"])
        valc = self._fwd(memory, a["val"], out_dtype=F32)
        if geo["pad_rows"] is not None:
            K.zero_masked_rows_(valc, geo["pad_rows"])
        val4 = valc.view(B, S, a["M"], -1)
        core = K.cast(K.msda_forward(val4, loc, attn, geo), T)
        z1 = self._fwd(core, a["out"], residual=t2)
        t1 = K.layernorm(z1, *self.ln_params(lw["ln1"]))
        h = self._fwd(t1, lw["l1"], relu=1)
        z3 = self._fwd(h, lw["l2"], residual=t1)
        t3 = K.layernorm(z3, *self.ln_params(lw["ln3"]))
        sv = dict(sine=sine, r1=r1, qk_in=qk_in, tgt=tgt, sactx=sactx, att=att, z2=z2, qca=qca, val4=val4, loc=loc, attn=attn,
                  core=core, z1=z1, t1=t1, h=h, z3=z3, ref=ref)
        return t3, sv

    def _dec_layer_bwd(self, lw, sv, d_t3, memory, dmem, geo, Qt, need_dmem):
        K, T = self.K, self.T
        B, S = geo["B"], geo["S"]
        a = lw["ca"]
        dz3_32, dz3 = self._ln_bwd(lw["ln3"], sv["z3"], d_t3)
        dh = self._bwd(lw["l2"], dz3, sv["h"], relu_mask=sv["h"])
        d_t1 = self._bwd(lw["l1"], dh, sv["t1"], residual=dz3_32, out_dtype=F32)
        del dh
        dz1_32, dz1 = self._ln_bwd(lw["ln1"], sv["z1"], d_t1)
        dcore = self._bwd(a["out"], dz1, sv["core"], out_dtype=F32)
        gv, gl, ga = K.msda_backward(sv["val4"], sv["loc"], sv["attn"], dcore.view(B, Qt, -1), geo)
        doa = K.msda_bwd_glue(gl, ga, sv["attn"], sv["ref"], geo["vr"], geo, Qt, a["M"], a["P"], T)
        dqca = self._bwd(a["oa"], doa, sv["qca"], out_dtype=F32)
        gv2 = gv.view(B * S, -1)
        if geo["pad_rows"] is not None:
            K.zero_masked_rows_(gv2, geo["pad_rows"])
        dval = K.cast(gv2, T)
        dmem = self._bwd(a["val"], dval, memory, need_dx=need_dmem, residual=dmem, out_dtype=F32)
        # t2 = LN2(z2) feeds the residual of the cross-attention block (dz1) and, through qca = t2 + qp, its query (dqca)
        dz2_32, dz2 = self._ln_bwd(lw["ln2"], sv["z2"], dz1_32, dqca)
        datt = self._bwd(lw["o"], dz2, sv["att"], out_dtype=T)
        dqk, dv = K.sa_backward(sv["sactx"], datt)
        dqk_in = self._bwd(lw["qk"], dqk, sv["qk_in"], out_dtype=F32)
        d_tgt = self._bwd(lw["v"], dv, sv["tgt"], residual=dz2_32, out_dtype=F32)
        d_tgt = K.add_cast(d_tgt, dqk_in, None, F32)
        if self.rph[0].train or self.rph[1].train:
            dqp = K.add_cast(dqca, dqk_in, None, T)                   # query_pos = ref_point_head(sine(ref)); ref is detached
            dr1 = self._bwd(self.rph[1], dqp, sv["r1"], relu_mask=sv["r1"])
            self._bwd(self.rph[0], dr1, sv["sine"], need_dx=False)
        return d_tgt, dmem

    # ------------------------------------------------------------------------------------------------ the step
    def forward_backward(self, samples, targets, loss_scale=1.0):
        """forward + loss_CTC + backward: fills the gradient arena (NOT zeroed here), returns the loss (0-d tensor)"""
        K, T, m, tr = self.K, self.T, self.model, self.model.transformer
        if hasattr(K, "begin"):
            K.begin(T)
        if isinstance(samples, (list, torch.Tensor)):
            samples = nested_tensor_from_tensor_list(samples)
        K.check_device(samples.tensors)              # train_ops: CUDA or raise (no CPU fallback)
        B = samples.tensors.shape[0]
        d = tr.d_model
        nopad = bool(getattr(samples, "nopad", False))
        live_seg = {e[3] for e in self.layout}
        train_front = any(n.startswith(("backbone.", "input_proj.")) for n, *_ in self.layout)
        train_enc = 1 in live_seg
        train_dec = 0 in live_seg or 2 in live_seg
        src_flatten = None
        if self.front is not None:
            src, pos32, mask_flatten, level_hw, masks = self.front.forward(samples)
        else:
            src_flatten, pos32, mask_flatten, level_hw, masks = self._front(samples)
        geo = self._geometry(B, level_hw, masks, mask_flatten, nopad)
        S = geo["S"]
        geo["ref_enc"] = K.enc_ref_points(geo["vr"], geo)
        pos = K.cast(pos32.reshape(B * S, d), T)
        if src_flatten is not None:
            src = K.cast(src_flatten.detach().reshape(B * S, d), T)

        # ---- encoder
        q = K.add(src, pos)
        enc_saved = []
        for i, lw in enumerate(self.enc):
            src, qn, sv = self._enc_layer_fwd(lw, src, q, pos, geo, i == len(self.enc) - 1)
            enc_saved.append(sv)
            q = qn
        memory = src

        # ---- queries: DN part (quirk Q3: targets are always passed) + learned table; anchors from the two-stage selection
        with torch.enable_grad():
            in_label, in_bbox, attn_mask, dn_meta = prepare_for_cdn(
                dn_args=(targets, m.dn_number, m.dn_label_noise_ratio, m.dn_box_noise_scale), training=True,
                num_queries=m.num_queries, num_classes=m.num_classes, hidden_dim=m.hidden_dim, label_enc=m.label_enc)
            Q = tr.num_queries
            tgt_tab = tr.tgt_embed.weight[None].expand(B, -1, -1)
            tgt_full = torch.cat([in_label, tgt_tab], 1) if in_label is not None and in_label.shape[1] > 0 else tgt_tab
        with torch.no_grad():
            ref0 = K.two_stage_refs(self, memory, geo)                                    # (B,Q,4) sigmoid space
            if in_bbox is not None and in_bbox.shape[1] > 0:
                ref_all = torch.cat([in_bbox.detach().sigmoid(), ref0], 1)
            else:
                ref_all, attn_mask = ref0, None
        Qt = ref_all.shape[1]
        ref = ref_all.reshape(B * Qt, 4).contiguous()
        tgt = K.cast(tgt_full.detach().reshape(B * Qt, d).contiguous(), T)
        mkey = (Qt, None if dn_meta is None else (dn_meta["pad_size"], dn_meta["num_dn_group"]))
        if attn_mask is None:
            mask_bool = None
        else:                               # the mask depends only on (pad_size, groups, Q): its bit-matrix form is built once per shape
            if mkey not in self._masks:
                if len(self._masks) > 64:
                    self._masks.clear()
                self._masks[mkey] = K.make_mask(attn_mask)
            mask_bool = self._masks[mkey]

        # ---- decoder
        dec_saved = []
        refs = [ref]
        for i, lw in enumerate(self.dec):
            tgt, sv = self._dec_layer_fwd(i, lw, tgt, ref, memory, mask_bool, geo, Qt)
            dec_saved.append(sv)
            ref = K.box_head(self, tgt, self.bbox[i], ref)
            refs.append(ref)
        n_dec = len(self.dec)
        hs = K.layernorm(tgt, *self.ln_params(tr.decoder.norm))
        logits = K.head_logits(hs, self.cls)                                               # fp32 (B*Qt, C), pitched
        boxes = K.box_head(self, hs, self.bbox[n_dec - 1], refs[n_dec - 1])
        C = logits.shape[1]

        # ---- loss_CTC forward + backward
        lens = [int(t["labels"].numel()) for t in targets]
        Lmax = max(lens) if lens else 0
        tt = torch.zeros((B, max(Lmax, 1)), dtype=torch.int32)
        for i, t in enumerate(targets):
            if lens[i]:
                tt[i, :lens[i]] = t["labels"].to(dtype=torch.int32, device="cpu") if not t["labels"].is_cuda else t["labels"].int().cpu()
        tt = tt.to(self.device, non_blocking=True)
        if Lmax == 0:
            tt = tt[:, :0]
        ll = torch.tensor(lens, dtype=torch.int32).to(self.device, non_blocking=True)
        loss, dlog = K.ctc_loss_grad(logits.view(B, Qt, C), boxes.view(B, Qt, 4), tt, ll, eps=0.003, zero_infinity=True)
        self.last = {"pred_logits": logits.view(B, Qt, C), "pred_boxes": boxes.view(B, Qt, 4), "dn_meta": dn_meta}
        if not train_dec and not train_enc and not train_front:
            return loss

        # ================================================================================================ backward
        scale = float(loss_scale) / float(self.world)
        pad = 8 if T != F32 else 4
        ldc = (C + pad - 1) // pad * pad
        dl16 = torch.zeros((B * Qt, ldc), dtype=T, device=self.device)
        dl16[:, :C] = dlog.view(B * Qt, C) * scale if scale != 1.0 else dlog.view(B * Qt, C)
        only_head = all(n.startswith(HEAD_PREFIXES) for n, *_ in self.layout)
        d_hs = self._bwd(self.cls, dl16[:, :C], hs, need_dx=not only_head, out_dtype=F32)
        if only_head:                       # step-1 fine-tuning (finetuning.py:531-539): only the class heads are optimised
            return loss
        d_t, _ = self._ln_bwd(tr.decoder.norm, tgt, d_hs, want16=False)
        dmem = None
        need_dmem = train_enc or train_front
        for i in range(n_dec - 1, -1, -1):
            d_t, dmem = self._dec_layer_bwd(self.dec[i], dec_saved[i], d_t, memory, dmem, geo, Qt, need_dmem)
            dec_saved[i] = None
        d_tgt0 = d_t
        self._segment_done(0)
        if need_dmem:
            lvl = None
            if tr.level_embed.requires_grad and self.grad(tr.level_embed) is not None:
                lvl = [self.grad(tr.level_embed)[l] for l in range(geo["nlev"])]
            d_y, d_qn = dmem, None
            dq = None
            for i in range(len(self.enc) - 1, -1, -1):
                d_y, dq = self._enc_layer_bwd(self.enc[i], enc_saved[i], d_y, d_qn, geo, train_front or i > 0, lvl)
                d_qn = dq
                enc_saved[i] = None
            self._segment_done(1)
            roots, grads = [tgt_full], [d_tgt0.view(B, Qt, d)]
            if train_front:
                dsrc0 = K.add_cast(d_y, dq, None, F32)                 # q0 = src0 + pos: both paths reach src_flatten
                if self.front is not None:
                    self.front.backward(dsrc0)
                else:
                    roots.append(src_flatten)
                    grads.append(dsrc0.view(B, S, d))
        else:
            roots, grads = [tgt_full], [d_tgt0.view(B, Qt, d)]
        roots_g = [(r, g) for r, g in zip(roots, grads) if r.requires_grad]
        if roots_g:
            torch.autograd.backward([r for r, _ in roots_g], [g for _, g in roots_g])
        self._segment_done(2)
        return loss

    # ------------------------------------------------------------------------------------------------ all-reduce / optimizer
    def _segment_done(self, seg):
        """N > 1: the gradients of arena segment `seg` are final (issued on the current stream) -> start their all-reduce"""
        if self.world <= 1:
            return
        import torch.distributed as dist
        for s, a, b in self.chunks:
            if s == seg and (a, b) not in self._reduced:
                self._works.append(dist.all_reduce(self.flat_g[a:b], group=self.pg, async_op=True))
                self._reduced.add((a, b))

    def zero_grad(self):
        self.flat_g.zero_()
        self._works, self._reduced = [], set()

    def optimizer_step(self):
        K = self.K
        if self.world > 1:
            for s, a, b in self.chunks:          # segments no backward touched in this step (e.g. frozen encoder): still consistent
                if (a, b) not in self._reduced:
                    self._segment_done(s)
            for w in self._works:
                w.wait()
        K.optim_begin(self.state)
        K.grad_sumsq(self.flat_g, self.state)
        for a, b, pg in self.group_ranges:
            K.adamw(self.flat_p[a:b], self.flat_g[a:b], self.flat_m[a:b], self.flat_v[a:b], pg["lr"], self.betas[0], self.betas[1], self.eps,
                    pg["weight_decay"], self.max_norm, self.state)
        self.repack()
        if getattr(self.model, "_engine", None) is not None:
            self.model.invalidate_engine()

    def step(self, samples, targets):
        """one fine-tune step (reference engine.py:192-241); returns the loss as a 0-d device tensor"""
        self.zero_grad()
        loss = self.forward_backward(samples, targets)
        self.optimizer_step()
        return loss

    def grad_norm(self):
        return float(torch.sqrt(self.state[0]))

    # ------------------------------------------------------------------------------------------------ checkpoint / resume
    def state_dict(self):
        """optimizer state for the reference's checkpoint dict (finetuning.py:664-674 saves 'optimizer'): both moments, the step count and
        the arena layout they refer to (parameter name -> offset, size), plus the current learning rates"""
        return {"step": float(self.state[1]), "exp_avg": self.flat_m.detach().cpu(), "exp_avg_sq": self.flat_v.detach().cpu(),
                "layout": [(n, o, k) for n, p, gi, seg, o, k in self.layout],
                "param_groups": [{"lr": g["lr"], "weight_decay": g["weight_decay"], "initial_lr": g["initial_lr"]} for g in self.param_groups]}

    def load_state_dict(self, sd):
        mine = {n: (o, k) for n, p, gi, seg, o, k in self.layout}
        m = sd["exp_avg"].to(self.device)
        v = sd["exp_avg_sq"].to(self.device)
        for n, o, k in sd["layout"]:           # by NAME: the arena order may differ (other param groups / frozen sets)
            if n in mine and mine[n][1] == k:
                self.flat_m[mine[n][0]:mine[n][0] + k].copy_(m[o:o + k])
                self.flat_v[mine[n][0]:mine[n][0] + k].copy_(v[o:o + k])
        self.state[1] = float(sd["step"])
        for g, s_ in zip(self.param_groups, sd.get("param_groups", [])):
            g.update(s_)

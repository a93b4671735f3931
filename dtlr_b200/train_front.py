"""Native front of the fine-tune step: ResNet-50 (frozen stem + layer1 forward only, layer2-4 forward AND backward) and
input_proj + GroupNorm forward and backward on libdtlr_b200 kernels -- the part of reference engine.py:192-241 that torch autograd runs
over cuDNN for models/dino/backbone.py:109-128 and models/dino/dino.py:118-135, 304-318.

Layout: NHWC rows [B*H*W, C] in the compute dtype.  FrozenBatchNorm (backbone.py:62-72) is folded into the operand copies of the
weights (w' = w * scale, bias' = bias - mean * scale); the gradient of the un-folded parameter is scale (x) dW', applied by the one
launch that also undoes the (tap, cin) operand order (dtlr_unpack_conv_grads).  Per convolution:
  1x1            forward dtlr_gemm; wgrad dtlr_wgrad; dgrad dtlr_gemm against the transposed copy
  3x3 stride 1   forward implicit GEMM (dtlr_conv2d_nhwc) when the shape allows, else dtlr_im2col + dtlr_gemm; wgrad dtlr_im2col +
                 dtlr_wgrad; dgrad = the same implicit-GEMM kernel over dY with the flipped-tap copy (or gemm + dtlr_col2im)
  3x3 stride 2, 1x1 stride 2 (downsample), input_proj[3]   dgrad dcol = dY . W' (dtlr_gemm) + dtlr_col2im (gather form of the transposed conv)
ReLU masks come from the saved outputs; the Bottleneck's closing ReLU hands its masked gradient on twice (fp32 along the identity path,
16-bit as conv3's operand: dtlr_relu_bwd_dual).
"""
import torch
import torch.nn.functional as F

from . import ops
from . import train_ops as K
from .misc import NestedTensor

F32 = torch.float32


class ConvW:
    """operand copies and gradient routing of one convolution"""

    def __init__(self, conv, bn, dtype, device, want_flipped=False):
        w = conv.weight
        self.param = w
        self.Cout, self.Cin, kh, kw = w.shape
        self.k, self.taps = kh, kh * kw
        self.stride, self.pad = conv.stride[0], conv.padding[0]
        self.train = w.requires_grad
        if bn is not None:
            scale, bias = bn.scale_bias()
            self.scale = scale.detach().float().contiguous()
            self.bias32 = bias.detach().float().contiguous()
            self.bias_param = None
        else:
            self.scale = None
            self.bias_param = conv.bias
            self.bias32 = None
        KK = self.taps * self.Cin
        pad = 8 if dtype != F32 else 4
        self.ldT = (self.Cout + pad - 1) // pad * pad
        self.fwd = torch.zeros((self.Cout, KK), dtype=dtype, device=device)
        self.bwdT = torch.zeros((KK, self.ldT), dtype=dtype, device=device)               # kind 1: plain transpose
        self.bwdF = torch.zeros((self.Cin, self.taps * self.Cout), dtype=dtype, device=device) if want_flipped else None
        self.gtmp = None          # fp32 scratch [Cout, taps*Cin] (set by NativeFront) or None: wgrad goes straight into the arena

    @property
    def bias(self):
        return self.bias32 if self.bias_param is None else self.bias_param.detach()


class NativeFront:
    def __init__(self, eng):
        from .engine import InferenceEngine
        self.eng = eng
        self.T = eng.T
        m = eng.model
        dev = eng.device
        body = m.backbone[0].body
        self.body = body
        half = self.T in ops.HALF
        # ---- frozen stem + layer1: packed once (these tensors never change during fine-tuning: requires_grad False, backbone.py:79-81)
        for p in list(body.conv1.parameters()) + list(body.layer1.parameters()):
            if p.requires_grad:
                raise NotImplementedError("NativeFront: conv1 / layer1 are expected frozen (reference backbone.py:79-81)")
        P = InferenceEngine(m).packed(self.T, dev)
        self.stem_gemm, self.stem_direct = P["stem_gemm"], P["stem_direct"]
        self.frozen_blocks = [b for b in P["blocks"] if b["layer"] == 1]
        self.return_layers = list(body.return_layers)
        # ---- trainable layer2-4
        self.blocks = []
        self.convs = []
        for li in (2, 3, 4):
            for blk in getattr(body, "layer%d" % li):
                d = {"layer": li, "stride": blk.stride,
                     "c1": self._conv(blk.conv1, blk.bn1), "c2": self._conv(blk.conv2, blk.bn2, flipped=half and blk.stride == 1),
                     "c3": self._conv(blk.conv3, blk.bn3), "ds": None}
                if blk.downsample is not None:
                    d["ds"] = self._conv(blk.downsample[0], blk.downsample[1])
                self.blocks.append(d)
        self.proj = []
        for seq in m.input_proj:
            conv, gn = seq[0], seq[1]
            self.proj.append({"conv": self._conv(conv, None), "gn": gn})
        # ---- scratch arena for the weight gradients that need the scale / layout fix-up, and the two device tables
        need = [c for c in self.convs if c.train and (c.scale is not None or c.taps > 1)]
        total = sum(c.Cout * c.taps * c.Cin for c in need)
        self.gscratch = torch.zeros(max(total, 1), dtype=F32, device=dev)
        off = 0
        for c in need:
            n = c.Cout * c.taps * c.Cin
            c.gtmp = self.gscratch[off:off + n].view(c.Cout, c.taps * c.Cin)
            off += n
        self._need = need
        rows, start = [], 0
        for c in self.convs:
            src = c.param.detach()
            rows.append([src.data_ptr(), c.Cout, c.Cin, c.taps, c.scale.data_ptr() if c.scale is not None else 0, c.fwd.data_ptr(),
                         c.bwdT.data_ptr(), 1, c.ldT, start])
            start += c.Cout * c.Cin * c.taps
            if c.bwdF is not None:      # a second pass over the same source for the flipped-tap copy (the forward copy is rewritten too)
                rows.append([src.data_ptr(), c.Cout, c.Cin, c.taps, c.scale.data_ptr() if c.scale is not None else 0, c.fwd.data_ptr(),
                             c.bwdF.data_ptr(), 2, c.taps * c.Cout, start])
                start += c.Cout * c.Cin * c.taps
        self.pack_table = torch.tensor(rows, dtype=torch.int64).to(dev)
        self.pack_n, self.pack_total = len(rows), start
        self._unpack = None
        self.repack()

    def _conv(self, conv, bn, flipped=False):
        c = ConvW(conv, bn, self.T, self.eng.device, flipped)
        self.convs.append(c)
        return c

    def repack(self):
        from . import _lib as L
        L.set_flavor(self.T)
        ops._call("dtlr_pack_conv", ops._p(self.pack_table), self.pack_n, _ll(self.pack_total), L._DT[self.T], ops._st(self.pack_table))

    def _unpack_tables(self):
        if self._unpack is None:
            rows, starts, s = [], [], 0
            for c in self._need:
                g = self.eng.grad(c.param)
                if g is None:
                    continue
                rows.append([c.gtmp.data_ptr(), g.data_ptr(), c.Cout, c.Cin, c.taps, c.scale.data_ptr() if c.scale is not None else 0])
                starts.append(s)
                s += c.Cout * c.Cin * c.taps
            dev = self.eng.device
            self._unpack = (torch.tensor(rows, dtype=torch.int64).to(dev) if rows else None,
                            torch.tensor(starts, dtype=torch.int64).to(dev) if rows else None, len(rows), s)
        return self._unpack

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, samples):
        """-> src (B*S, d) in the compute dtype, pos fp32 (B,S,d) incl. level_embed, mask_flatten, level_hw, masks"""
        eng, T = self.eng, self.T
        m, tr = eng.model, eng.model.transformer
        x = samples.tensors.float().contiguous()
        B, _, H, W = x.shape
        d = tr.d_model
        # ---- frozen part (inference kernels)
        if self.stem_gemm is not None and ops.STEM_TENSOR_CORE:
            col, Ho, Wo = ops.im2col(x, B, H, W, 3, 7, 7, 2, 3, T, nchw_input=True, ldo=self.stem_gemm[0].shape[1])
            y = ops.gemm(col, *self.stem_gemm, relu=1)
        else:
            y, Ho, Wo = ops.stem_conv(x, *self.stem_direct, B, H, W, T)
        y, Hc, Wc = ops.maxpool3x3s2(y, B, Ho, Wo, 64)
        cin = 64
        for blk in self.frozen_blocks:
            a = ops.gemm(y, *blk["c1"], relu=1)
            planes = blk["c1"][0].shape[0]
            if ops.conv2d_nhwc_supported(a, Hc, Wc, planes, 3, 1):
                bm, _, _ = ops.conv2d_nhwc(a, *blk["c2"], B, Hc, Wc, planes, 3, 1, relu=1)
            else:
                bm = ops.gemm(ops.im2col(a, B, Hc, Wc, planes, 3, 3, 1, 1, T)[0], *blk["c2"], relu=1)
            idt = ops.gemm(y, *blk["ds"]) if blk["ds"] is not None else y
            y = ops.gemm(bm, *blk["c3"], residual=idt, relu=2)
            cin = planes * 4
        # ---- layer2-4 with everything the backward needs
        saved, feats = [], []
        for i, blk in enumerate(self.blocks):
            s = blk["stride"]
            c1, c2, c3, ds = blk["c1"], blk["c2"], blk["c3"], blk["ds"]
            a = ops.gemm(y, c1.fwd, c1.bias, relu=1)
            p = c1.Cout
            col = None
            if ops.conv2d_nhwc_supported(a, Hc, Wc, p, 3, s):
                bm, Hn, Wn = ops.conv2d_nhwc(a, c2.fwd, c2.bias, B, Hc, Wc, p, 3, 1, relu=1, stride=s)
            else:
                col, Hn, Wn = ops.im2col(a, B, Hc, Wc, p, 3, 3, s, 1, T)
                bm = ops.gemm(col, c2.fwd, c2.bias, relu=1)
            xs = None
            if ds is not None:
                if s == 1:
                    idt = ops.gemm(y, ds.fwd, ds.bias)
                else:
                    xs = ops.im2col(y, B, Hc, Wc, cin, 1, 1, s, 0, T)[0]          # the strided rows (also conv's wgrad operand)
                    idt = ops.gemm(xs, ds.fwd, ds.bias)
            else:
                idt = y
            yo = ops.gemm(bm, c3.fwd, c3.bias, residual=idt, relu=2)
            saved.append(dict(x=y, a=a, bm=bm, y=yo, xs=xs, col=col, Hin=Hc, Win=Wc, Hout=Hn, Wout=Wn, cin=cin))
            y, Hc, Wc, cin = yo, Hn, Wn, p * 4
            last = i == len(self.blocks) - 1 or self.blocks[i + 1]["layer"] != blk["layer"]
            if last and blk["layer"] in self.return_layers:
                feats.append((y, Hc, Wc, cin, i))
        # ---- level geometry, masks, positions (torch index bookkeeping on small tensors, no gradient)
        nlev = len(self.proj)
        level_hw = [(h, w) for _, h, w, _, _ in feats]
        for l in range(len(feats), nlev):
            c = self.proj[l]["conv"]
            h_in, w_in = level_hw[-1]
            level_hw.append(((h_in + 2 * c.pad - c.k) // c.stride + 1, (w_in + 2 * c.pad - c.k) // c.stride + 1))
        S = sum(h * w for h, w in level_hw)
        starts = [0]
        for h, w in level_hw[:-1]:
            starts.append(starts[-1] + h * w)
        with torch.no_grad():
            masks = [F.interpolate(samples.mask[None].float(), size=hw).to(torch.bool)[0] for hw in level_hw]
            pe = m.backbone[1]
            pos = torch.cat([pe(NestedTensor(None, mk)).flatten(2).transpose(1, 2) + tr.level_embed.detach()[l].view(1, 1, -1)
                             for l, mk in enumerate(masks)], 1)
            mask_flatten = torch.cat([mk.flatten(1) for mk in masks], 1)
        # ---- input_proj + GroupNorm straight into the token tensor
        src = torch.empty((B * S, d), dtype=T, device=x.device)
        psaved = []
        for l in range(nlev):
            pj = self.proj[l]
            c, gn = pj["conv"], pj["gn"]
            h, w = level_hw[l]
            if l < len(feats):
                f = feats[l][0]
                a_in = f
            else:
                assert l == len(feats), "one extra level (num_feature_levels = backbone levels + 1: every shipped config)"
                f, fh, fw, fc, _ = feats[-1]
                a_in = ops.im2col(f, B, fh, fw, fc, c.k, c.k, c.stride, c.pad, T)[0]
            y32 = ops.gemm(a_in, c.fwd, c.bias, out_dtype=F32)
            ops.groupnorm_into(y32, gn.weight.detach(), gn.bias.detach(), src, B, h * w, d, gn.num_groups, starts[l], S, eps=gn.eps)
            psaved.append(dict(a_in=a_in, y32=y32))
        self._saved = dict(blocks=saved, feats=feats, proj=psaved, level_hw=level_hw, starts=starts, B=B, S=S)
        return src, pos, mask_flatten, level_hw, masks

    # ------------------------------------------------------------------------------------------------ backward
    def _wgrad(self, c, dy, x):
        if not c.train:
            return
        if c.gtmp is not None:
            K.wgrad(dy, x, c.gtmp)
        else:
            K.wgrad(dy, x, self.eng.grad(c.param).view(c.Cout, c.Cin))

    def backward(self, dsrc32):
        """dsrc32 fp32 (B*S, d): gradient of the loss with respect to src_flatten.  Fills the arena gradients of input_proj and layer2-4."""
        eng, T = self.eng, self.T
        sv = self._saved
        B, S, level_hw, starts = sv["B"], sv["S"], sv["level_hw"], sv["starts"]
        d = dsrc32.shape[1]
        feats = sv["feats"]
        self.gscratch.zero_()
        dfeat = [None] * len(feats)
        for l in range(len(self.proj) - 1, -1, -1):
            pj, ps = self.proj[l], sv["proj"][l]
            c, gn = pj["conv"], pj["gn"]
            h, w = level_hw[l]
            gg = eng.grad(gn.weight) if gn.weight.requires_grad else None
            gb = eng.grad(gn.bias) if gn.bias.requires_grad else None
            dy = torch.empty((B * h * w, d), dtype=T, device=dsrc32.device)
            ops._call("dtlr_groupnorm_bwd", ops._p(ps["y32"]), ops._p(dsrc32[starts[l]:]), _ll(S), ops._p(gn.weight.detach()), ops._p(dy),
                      ops._p(gg), ops._p(gb), B, h * w, d, gn.num_groups, _cf(gn.eps), _code(dy), ops._st(dy))
            if c.bias_param is not None and c.bias_param.requires_grad and eng.grad(c.bias_param) is not None:
                K.colsum(dy, eng.grad(c.bias_param))
            self._wgrad(c, dy, ps["a_in"])
            if l < len(feats):
                dfeat[l] = K.gemm(dy, c.bwdT[:, :c.Cout], None, residual=dfeat[l], out_dtype=F32)
            else:
                f, fh, fw, fc, _ = feats[-1]
                dcol = K.gemm(dy, c.bwdT[:, :c.Cout], None, out_dtype=F32)
                if dfeat[-1] is None:
                    dfeat[-1] = torch.empty((B * fh * fw, fc), dtype=F32, device=dsrc32.device)
                    acc = 0
                else:
                    acc = 1
                _col2im(dcol, dfeat[-1], B, fh, fw, fc, c.k, c.stride, c.pad, h, w, acc)
        # ---- layer4 .. layer2
        feat_of_block = {bi: l for l, (_, _, _, _, bi) in enumerate(feats)}
        dy = None
        for i in range(len(self.blocks) - 1, -1, -1):
            blk, s_ = self.blocks[i], sv["blocks"][i]
            if i in feat_of_block:
                g = dfeat[feat_of_block[i]]
                dy = g if dy is None else K.add_cast(dy, g, None, F32)
            if dy is None:
                continue
            dy = self._block_bwd(blk, s_, dy, B, need_dx=i > 0)
            sv["blocks"][i] = None
        tab, st, n, total = self._unpack_tables()
        if n:
            ops._call("dtlr_unpack_conv_grads", ops._p(tab), ops._p(st), n, _ll(total), ops._st(tab))
        self._saved = None

    def _block_bwd(self, blk, sv, dy32, B, need_dx):
        T = self.T
        c1, c2, c3, ds = blk["c1"], blk["c2"], blk["c3"], blk["ds"]
        s = blk["stride"]
        Hi, Wi, Ho, Wo, cin = sv["Hin"], sv["Win"], sv["Hout"], sv["Wout"], sv["cin"]
        p = c1.Cout
        dsum = torch.empty(sv["y"].shape, dtype=T, device=dy32.device)
        ops._call("dtlr_relu_bwd_dual", ops._p(dy32), ops._p(sv["y"]), ops._p(dsum), _ll(dy32.numel()), _code(dsum), ops._st(dsum))
        # conv3 (1x1)
        self._wgrad(c3, dsum, sv["bm"])
        db = K.gemm(dsum, c3.bwdT[:, :c3.Cout], None, residual=sv["bm"], relu=3, out_dtype=T)      # ReLU mask of bm in the epilogue
        # conv2 (3x3, stride s)
        col = sv["col"] if sv["col"] is not None else ops.im2col(sv["a"], B, Hi, Wi, p, 3, 3, s, 1, T)[0]
        self._wgrad(c2, db, col)
        del col
        if c2.bwdF is not None and ops.conv2d_nhwc_supported(db, Ho, Wo, p, 3, 1):
            da = ops.conv2d_nhwc(db, c2.bwdF, None, B, Ho, Wo, p, 3, 1, relu=3, residual=sv["a"])[0]      # masked by a > 0
        else:
            dcol = K.gemm(db, c2.bwdT[:, :c2.Cout], None, out_dtype=F32)
            da32 = torch.empty((B * Hi * Wi, p), dtype=F32, device=dy32.device)
            _col2im(dcol, da32, B, Hi, Wi, p, 3, s, 1, Ho, Wo, 0)
            da = K.cast(da32, T)
            K.relu_bwd_(da, sv["a"])
        # conv1 (1x1) and the identity path
        self._wgrad(c1, da, sv["x"])
        if ds is None:
            return K.gemm(da, c1.bwdT[:, :c1.Cout], None, residual=dy32, out_dtype=F32) if need_dx else None
        self._wgrad(ds, dsum, sv["xs"] if sv["xs"] is not None else sv["x"])
        if not need_dx:
            return None
        dxs = K.gemm(dsum, ds.bwdT[:, :ds.Cout], None, out_dtype=F32)
        if s == 1:
            return K.gemm(da, c1.bwdT[:, :c1.Cout], None, residual=dxs, out_dtype=F32)
        dx = K.gemm(da, c1.bwdT[:, :c1.Cout], None, out_dtype=F32)
        _col2im(dxs, dx, B, Hi, Wi, cin, 1, s, 0, Ho, Wo, 1)
        return dx


def _ll(v):
    import ctypes
    return ctypes.c_longlong(int(v))


def _cf(v):
    import ctypes
    return ctypes.c_float(float(v))


def _code(t):
    from . import _lib as L
    return L.dtype_code(t)


def _col2im(dcol, dx, B, H, W, C, k, stride, pad, Ho, Wo, accumulate):
    assert dcol.dtype == F32 and dx.dtype == F32 and dcol.stride(1) == 1 and dx.is_contiguous()
    ops._call("dtlr_col2im", ops._p(dcol), dcol.stride(0), ops._p(dx), B, H, W, C, k, k, stride, pad, Ho, Wo, int(accumulate), ops._st(dx))

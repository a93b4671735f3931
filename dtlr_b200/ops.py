"""Python wrappers (torch tensors in, torch tensors out) over the C ABI of include/dtlr_b200.h.
PyTorch only provides device memory and the current stream here; every computation is a libdtlr_b200 kernel."""
import torch

from . import _lib as L

HALF = (torch.bfloat16, torch.float16)      # the two 16-bit flavours of the throughput mode (one library each, _lib.set_flavor)



class _SplitOut:
    """out_dtype marker of gemm / conv2d_nhwc: write the fp32 result as the split operand [hi | hi | lo] (include/dtlr_b200.h DTLR_SPLIT16)"""

    def __repr__(self):
        return "ops.SPLIT"


SPLIT = _SplitOut()


def gemm(a, w, bias=None, residual=None, relu=False, out_dtype=None, out=None, split3=False):
    """C = act(a @ w.T + bias) (+ residual); relu: 0/False none, 1/True before the residual add, 2 after it.  a (M,K) row-major (last-dim stride 1, any row pitch), w (N,K),
    bias fp32 (N) or None, residual (M,N) of the output dtype or None.
    Split-precision products: an fp32 `a` against a 16-bit `w` of 3K columns ([hi | lo | hi], engine._split_w) is split on the
    fly ([hi | hi | lo], split_cast) and multiplied on the 16-bit tensor-core kernel with an fp32 result; out_dtype=SPLIT stores that result as
    the split operand (M, 3N) of the next product; split3=True marks 16-bit operands that already ARE whole-row split matrices (the tile kernel
    then fetches each hi / lo tile once)."""
    L.require_cuda(a, w, bias, residual)
    k_alg = a.shape[1]                     # algorithmic K (the split product runs 3K columns for it)
    if a.dtype == torch.float32 and w.dtype in HALF:
        # split-precision product (tensor-core parity mode): w is a weight packed [hi | lo | hi] (engine._split_w),
        # a becomes [hi | hi | lo] (dtlr_split_cast); the 16-bit tcgen05 GEMM over K' = 3K accumulates hi.hi + hi.lo + lo.hi in fp32
        assert a.dim() == 2 and w.dim() == 2 and w.shape[1] == 3 * a.shape[1], (a.shape, w.shape)
        if out_dtype is None:
            out_dtype = torch.float32
        a = split_cast(a, w.dtype)
        split3 = True
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1], (a.shape, w.shape)
    assert a.stride(1) == 1 and w.stride(1) == 1 and a.dtype == w.dtype
    M, K = a.shape
    N = w.shape[0]
    if out_dtype is None:
        out_dtype = a.dtype
    split_out = out_dtype is SPLIT        # fp32 result stored as the 16-bit [hi | hi | lo] operand of the next split product: (M, 3N)
    if split_out:
        assert a.dtype in HALF and out is None and N % 4 == 0
        out = torch.empty((M, 3 * N), dtype=a.dtype, device=a.device)
        out_dtype = torch.float32         # (dtype of the residual)
    elif out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert split_out or (out.stride(1) == 1 and out.shape == (M, N) and out.dtype == out_dtype)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
    ldr = 0
    if residual is not None:
        assert residual.dtype == out_dtype and residual.shape == (M, N) and residual.stride(1) == 1
        ldr = residual.stride(0)
    if L.TIMER is not None:
        L.TIMER("gemm", 2.0 * M * N * k_alg, a.device, True)
        L.GEMM_BYTES += (M * K * a.element_size() + N * K * w.element_size() + M * N * out.element_size()
                         + (M * N * residual.element_size() if residual is not None else 0))
    in_code, out_code = L.dtype_code(a), L.dtype_code(out)     # (16-bit tensors select the library flavour before L.lib())
    if split_out:
        out_code = L.SPLIT16
    if split3 and SPLIT3_LOADS:    # a = [hi | hi | lo], w = [hi | lo | hi] over whole rows: the tile kernel loads each hi / lo tile once (DTLR_SPLIT16 in)
        assert a.dtype in HALF and K % 3 == 0
        in_code = L.SPLIT16
    with torch.cuda.device(a.device):
        rc = L.lib().dtlr_gemm(L.ptr(a), a.stride(0), L.ptr(w), w.stride(0), L.ptr(bias) if bias is not None else None,
                               L.ptr(residual) if residual is not None else None, ldr, L.ptr(out), out.stride(0),
                               M, N, K, in_code, out_code, int(relu), L.stream_ptr(a.device))
    L.check(rc, "dtlr_gemm")
    if L.TIMER is not None:
        L.TIMER("gemm", 0.0, a.device, False)
    return out


# ------------------------------------------------------------------------------------------------------------------
import os as _os
LN_FUSE_MIN_K = int(_os.environ.get("DTLR_LN_FUSE_MIN_K", "1024"))
LN_FUSE_WS = _os.environ.get("DTLR_LN_FUSE_WS", "1") != "0"
STEM_TENSOR_CORE = _os.environ.get("DTLR_STEM_TC", "1") != "0"


def _call(name, *args):
    L.check(getattr(L.lib(), name)(*args), name)


def _p(t):
    return L.ptr(t) if t is not None else None


def _st(t):
    return L.stream_ptr(t.device)


def split_cast(x, dtype=torch.float16):
    """fp32 (M,K) (last-dim stride 1, any row pitch) -> 16-bit (M,3K) = [hi | hi | lo]: the A operand of a split-precision product."""
    import ctypes
    L.require_cuda(x)
    assert x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1 and dtype in HALF
    M, K = x.shape
    L.set_flavor(dtype)
    out = torch.empty((M, 3 * K), dtype=dtype, device=x.device)
    _call("dtlr_split_cast", _p(x), ctypes.c_longlong(x.stride(0)), _p(out), ctypes.c_longlong(M), K, _st(x))
    return out


def im2col(x, B, H, W, C, KH, KW, stride, pad, out_dtype, nchw_input=False, ldo=None):
    Ho = (H + 2 * pad - KH) // stride + 1
    Wo = (W + 2 * pad - KW) // stride + 1
    K = KH * KW * C
    ldo = ldo or K
    out = torch.empty((B * Ho * Wo, ldo), dtype=out_dtype, device=x.device)
    _call("dtlr_im2col", _p(x), _p(out), B, H, W, C, KH, KW, stride, pad, Ho, Wo, ldo, L.dtype_code(x),
          L.dtype_code(out), 1 if nchw_input else 0, _st(x))
    return out, Ho, Wo


CONV_STRIDED_IMPLICIT = _os.environ.get("DTLR_CONV_STRIDED_IMPLICIT", "1") != "0"
# split-precision mode: 3x3 convs as the implicit GEMM over 3C-channel pixels with an fp32 result (0: 16-bit im2col + GEMM, A/B)
SPLIT_CONV_IMPLICIT = _os.environ.get("DTLR_SPLIT_CONV_IMPLICIT", "1") != "0"
# split-precision mode: producers whose result only feeds another contraction (conv1 -> conv2 -> conv3 of a bottleneck, linear1 of an FFN,
# MLP hidden layers) write the split operand from their epilogue (DTLR_SPLIT16) instead of fp32 + dtlr_split_cast (0: A/B)
SPLIT_OUT_FUSED = _os.environ.get("DTLR_SPLIT_OUT_FUSED", "1") != "0"
# split-precision products tell dtlr_gemm that their operands are [hi | hi | lo] x [hi | lo | hi] (in_dtype DTLR_SPLIT16): the tile kernel
# then fetches A_hi, A_lo, W_hi, W_lo once per logical k-block instead of six tiles (0: plain walk over 3K, A/B)
SPLIT3_LOADS = _os.environ.get("DTLR_SPLIT3_LOADS", "1") != "0"


def conv2d_nhwc_supported(x, H, W, C, k, stride):
    """implicit-GEMM conv (csrc/gemm.cu): 16-bit NHWC, 'same' padding k = 2*pad+1, stride 1 or 2, C % 64 == 0 and an OUTPUT width that
    tiles into 128-pixel row segments"""
    if stride not in (1, 2) or (stride == 2 and not CONV_STRIDED_IMPLICIT):
        return False
    Wo = (W + 2 * (k // 2) - k) // stride + 1
    seg = 128 if Wo >= 128 else Wo
    return (x.dtype in HALF and C % 64 == 0 and seg >= 8 and 128 % seg == 0 and Wo % seg == 0)


def conv2d_nhwc(x, w, bias, B, H, W, C, k, pad, relu=0, residual=None, stride=1, out_dtype=None):
    """'same'-padded k x k conv with stride 1 or 2 as implicit GEMM: x 16-bit [B*H*W, C] NHWC, w 16-bit [Cout, k*k*C] ->
    16-bit (or fp32: out_dtype, used by the split-precision mode with C = 3 x channels) [B*Ho*Wo, Cout]; returns (out, Ho, Wo)"""
    Cout = w.shape[0]
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    if out_dtype is SPLIT:
        out = torch.empty((B * Ho * Wo, 3 * Cout), dtype=x.dtype, device=x.device)
        code = L.SPLIT16
    else:
        out = torch.empty((B * Ho * Wo, Cout), dtype=out_dtype or x.dtype, device=x.device)
        code = L.dtype_code(out)
    _call("dtlr_conv2d_nhwc_strided", _p(x), _p(w), _p(bias), _p(residual), _p(out), B, H, W, C, Cout, k, k, pad, int(stride), int(relu),
          code, _st(x))
    return out, Ho, Wo


def stem_conv(x, w_khkwcico, bias, B, H, W, out_dtype):
    Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    out = torch.empty((B * Ho * Wo, 64), dtype=out_dtype, device=x.device)
    _call("dtlr_stem_conv", _p(x), _p(w_khkwcico), _p(bias), _p(out), B, H, W, Ho, Wo, L.dtype_code(out), _st(x))
    return out, Ho, Wo


def maxpool3x3s2(x, B, H, W, C):
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    out = torch.empty((B * Ho * Wo, C), dtype=x.dtype, device=x.device)
    _call("dtlr_maxpool3x3s2", _p(x), _p(out), B, H, W, C, Ho, Wo, L.dtype_code(x), _st(x))
    return out, Ho, Wo


def groupnorm_into(x_f32, gamma, beta, out, B, HW, C, G, row_offset, rows_per_batch, eps=1e-5):
    """x_f32 [B*HW, C] fp32 -> out[b, row_offset + hw, :] of the (B, rows_per_batch, C) token buffer."""
    import ctypes
    dst = out.view(-1, C)[row_offset:]
    _call("dtlr_groupnorm", _p(x_f32), _p(gamma), _p(beta), _p(dst), B, HW, C, G, ctypes.c_longlong(rows_per_batch),
          ctypes.c_float(eps), L.dtype_code(out), _st(out))


def pos_sine_into(mask_u8, level_embed, out, B, H, W, npf, temp_h, temp_w, row_offset, rows_per_batch):
    import ctypes
    C = 2 * npf
    dst = out.view(-1, C)[row_offset:]
    _call("dtlr_pos_sine", _p(mask_u8), _p(level_embed), _p(dst), B, H, W, npf, ctypes.c_float(temp_h),
          ctypes.c_float(temp_w), ctypes.c_longlong(rows_per_batch), L.dtype_code(out), _st(out))


def add_layernorm(x, res, gamma, beta, add2=None, eps=1e-5, out=None):
    import ctypes
    rows, C = x.shape
    y = torch.empty_like(x) if out is None else out
    assert y.is_contiguous() and y.shape == x.shape
    y2 = torch.empty_like(x) if add2 is not None else None
    _call("dtlr_add_layernorm", _p(x), _p(res), _p(gamma), _p(beta), _p(y), _p(add2), _p(y2), ctypes.c_longlong(rows), C,
          ctypes.c_float(eps), L.dtype_code(x), _st(x))
    return (y, y2) if add2 is not None else y


def add(a, b):
    import ctypes
    out = torch.empty_like(a)
    _call("dtlr_add", _p(a), _p(b), _p(out), ctypes.c_longlong(a.numel()), L.dtype_code(a), _st(a))
    return out


def zero_masked_rows_(x, rowmask_u8):
    import ctypes
    _call("dtlr_zero_masked_rows", _p(x), _p(rowmask_u8), ctypes.c_longlong(x.shape[0]), x.shape[1], L.dtype_code(x), _st(x))
    return x


def msda_prep(proj_f32, ref, valid_ratios, shapes_host, n_levels, B, Lq, M, P):
    loc = torch.empty((B, Lq, M, n_levels, P, 2), dtype=torch.float32, device=proj_f32.device)
    attn = torch.empty((B, Lq, M, n_levels, P), dtype=torch.float32, device=proj_f32.device)
    _call("dtlr_msda_prep", _p(proj_f32), proj_f32.stride(0), _p(ref), ref.shape[-1], _p(valid_ratios), shapes_host, n_levels,
          _p(loc), _p(attn), B, Lq, M, P, _st(proj_f32))
    return loc, attn


def enc_ref_points(valid_ratios, shapes_host, n_levels, B, S):
    ref = torch.empty((B * S, 2), dtype=torch.float32, device=valid_ratios.device)
    _call("dtlr_enc_ref_points", _p(valid_ratios), shapes_host, n_levels, _p(ref), B, S, _st(valid_ratios))
    return ref


def encoder_proposals(memory, pad_u8, valid_hw_i32, shapes_host, n_levels, B, S, C, default_hw):
    import ctypes
    out_mem = torch.empty_like(memory)
    prop = torch.empty((B * S, 4), dtype=torch.float32, device=memory.device)
    _call("dtlr_encoder_proposals", _p(memory), _p(pad_u8), _p(valid_hw_i32), shapes_host, n_levels, _p(out_mem), _p(prop),
          B, S, C, ctypes.c_float(default_hw), L.dtype_code(memory), _st(memory))
    return out_mem, prop


def rowmax(x_f32, N):
    import ctypes
    out = torch.empty((x_f32.shape[0],), dtype=torch.float32, device=x_f32.device)
    _call("dtlr_rowmax", _p(x_f32), x_f32.stride(0), N, _p(out), ctypes.c_longlong(x_f32.shape[0]), _st(x_f32))
    return out


def sine_embed(ref, valid_ratios, B, Q, n_levels, out_dtype):
    out = torch.empty((B * Q, 512), dtype=out_dtype, device=ref.device)
    _call("dtlr_sine_embed", _p(ref), _p(valid_ratios), _p(out), B, Q, n_levels, L.dtype_code(out), _st(ref))
    return out


def box_refine(delta_f32, ref):
    import ctypes
    out = torch.empty_like(ref)
    _call("dtlr_box_refine", _p(delta_f32), delta_f32.stride(0), _p(ref), _p(out), ctypes.c_longlong(ref.shape[0]), _st(ref))
    return out


def sigmoid(x):
    import ctypes
    out = torch.empty_like(x)
    _call("dtlr_sigmoid", _p(x), _p(out), ctypes.c_longlong(x.numel()), _st(x))
    return out


def cast(x, dtype):
    import ctypes
    if x.dtype == dtype:
        return x
    out = torch.empty(x.shape, dtype=dtype, device=x.device)
    _call("dtlr_cast", _p(x), _p(out), ctypes.c_longlong(x.numel()), L.dtype_code(x), L.dtype_code(out), _st(x))
    return out


# decoder self-attention without a mask (inference): "tc" (default) = the single-pass tcgen05 / TMEM kernel (206 us per layer at
# B = 64, Q = 900: P in tensor memory, S and P.V issued by two warps that poll the four query-tile pipelines); "hmma" = the mma.sync
# flash kernel (232 us; also the path for masks, other head sizes and Q > 1024).  Both are bound by the exponentials (MUFU) and the
# softmax ALU work, not by the tensor pipe -- DESIGN.md 3.3
ATTN_IMPL = _os.environ.get("DTLR_ATTN", "tc")


def mha_self_attention(qk, k_off, v, attn_mask_u8, B, Q, heads, head_dim):
    L.set_flavor(v.dtype)
    out = torch.empty((B * Q, heads * head_dim), dtype=v.dtype, device=v.device)
    if ATTN_IMPL == "tc" and v.dtype in HALF and attn_mask_u8 is None and head_dim == 32 and 0 < Q <= 1024:
        vt = torch.empty((B * heads * 32, 1024), dtype=v.dtype, device=v.device)
        rc = L.lib().dtlr_mha_tcgen05(_p(qk), qk.stride(0), k_off, _p(v), v.stride(0), _p(vt), _p(out), out.stride(0), B, Q, heads,
                                      head_dim, _st(v))
        if rc == 0:
            L.LAUNCHES += 2
            return out
        if rc != 3:
            L.check(rc, "dtlr_mha_tcgen05")
    _call("dtlr_mha_self_attention", _p(qk), qk.stride(0), k_off, _p(v), v.stride(0), _p(attn_mask_u8), _p(out), out.stride(0),
          B, Q, heads, head_dim, L.dtype_code(v), _st(v))
    return out


def ctc_decode(pred_logits, pred_boxes, eps=0.003, want_new_pred=False, prob_scale=1.0):
    """fused CTC-view decode: pred_logits fp32 (B,Q,C), pred_boxes fp32 (B,Q,4) -> frames int32 (B,Q) in reading order
    (0 = blank, c+1 = class c) [, new_pred_logits fp32 (B,Q,C+1)].  prob_scale multiplies every class probability before the
    blank synthesis (reference ngram/prediction_helpers.py:5-46 `multiply_pred_logits_by`)."""
    import ctypes
    L.require_cuda(pred_logits, pred_boxes)
    B, Q, C = pred_logits.shape
    logits = pred_logits.float()
    if not (logits.stride(2) == 1 and logits.stride(0) == Q * logits.stride(1)):      # rows may be pitched (engine: 166 -> 168)
        logits = logits.contiguous()
    boxes = pred_boxes.float().contiguous()
    dev = logits.device
    frames = torch.empty((B, Q), dtype=torch.int32, device=dev)
    label = torch.empty((B, Q), dtype=torch.int32, device=dev)
    perm = torch.empty((B, Q), dtype=torch.int32, device=dev) if want_new_pred else None
    rsum = torch.empty((B, Q), dtype=torch.float32, device=dev) if want_new_pred else None
    newp = torch.empty((B, Q, C + 1), dtype=torch.float32, device=dev) if want_new_pred else None
    _call("dtlr_ctc_decode_scaled", _p(logits), logits.stride(1), _p(boxes), _p(frames), _p(perm), _p(newp), _p(label), _p(rsum),
          B, Q, C, ctypes.c_float(eps), ctypes.c_float(prob_scale), _st(logits))
    return (frames, newp) if want_new_pred else frames


def topk_select(scores, K):
    """scores fp32 (B,S) -> int64 (B,K) = torch.topk(scores, K, dim=1)[1] (csrc/select.cu; ties -> lowest index)"""
    L.require_cuda(scores)
    B, S = scores.shape
    scores = scores.contiguous()
    idx = torch.empty((B, K), dtype=torch.int64, device=scores.device)
    _call("dtlr_topk_select", _p(scores), B, S, K, _p(idx), _st(scores))
    return idx


def select_gather(idx, coord, prop, mem):
    """idx int64 (B,K); coord (bbox-head output) / prop (anchors, logit space) fp32 (B,S,4); mem (B,S,d) ->
    sigmoid(coord + prop)[idx] (B,K,4), sigmoid(prop)[idx] (B,K,4), mem[idx] (B,K,d)"""
    L.require_cuda(idx, coord, prop, mem)
    B, K = idx.shape
    S, d = mem.shape[1], mem.shape[2]
    assert coord.is_contiguous() and prop.is_contiguous() and mem.is_contiguous() and idx.is_contiguous()
    ref = torch.empty((B, K, 4), dtype=torch.float32, device=mem.device)
    box = torch.empty((B, K, 4), dtype=torch.float32, device=mem.device)
    tgt = torch.empty((B, K, d), dtype=mem.dtype, device=mem.device)
    _call("dtlr_select_gather", _p(idx), _p(coord), _p(prop), _p(mem), _p(ref), _p(box), _p(tgt), B, S, K, d, L.dtype_code(mem), _st(mem))
    return ref, box, tgt


def postprocess(pred_logits, pred_boxes, target_sizes, num_select, box_mode=0, nms_iou=None, score_thr=0.0):
    """PostProcess of reference dino.py:1008-1046 as kernels (csrc/select.cu): returns scores (B,K) fp32, labels (B,K) int64, boxes
    (B,K,4) fp32 [, keep (B,K) bool, read_labels (B,K) int32, read_count (B) int32 when nms_iou is not None]."""
    import ctypes
    L.require_cuda(pred_logits, pred_boxes, target_sizes)
    B, Q, C = pred_logits.shape
    logits = pred_logits.float()
    if not (logits.stride(2) == 1 and logits.stride(0) == Q * logits.stride(1)):
        logits = logits.contiguous()
    boxes = pred_boxes.float().contiguous()
    sizes = target_sizes.float().contiguous()
    dev = logits.device
    K = int(num_select)
    scores = torch.empty((B, K), dtype=torch.float32, device=dev)
    labels = torch.empty((B, K), dtype=torch.int32, device=dev)
    out_boxes = torch.empty((B, K, 4), dtype=torch.float32, device=dev)
    keep = rl = rc = None
    if nms_iou is not None:
        keep = torch.empty((B, K), dtype=torch.uint8, device=dev)
        rl = torch.empty((B, K), dtype=torch.int32, device=dev)
        rc = torch.empty((B,), dtype=torch.int32, device=dev)
    _call("dtlr_postprocess", _p(logits), logits.stride(1), _p(boxes), _p(sizes), B, Q, C, K, int(box_mode),
          ctypes.c_float(nms_iou if nms_iou is not None else -1.0), ctypes.c_float(score_thr), _p(scores), _p(labels), _p(out_boxes),
          _p(keep), _p(rl), _p(rc), _st(logits))
    if nms_iou is None:
        return scores, labels.long(), out_boxes
    L.LAUNCHES += 1
    return scores, labels.long(), out_boxes, keep.bool(), rl, rc


class _FusedCTCLoss(torch.autograd.Function):
    """loss_CTC of reference models/dino/dino.py:457-551 as ONE autograd node over the fused kernels of csrc/decode.cu
    (dtlr_ctc_loss): forward computes the loss AND d loss / d pred_logits (alpha-beta over the implicit interleaved-blank
    sequence); backward scales it.  pred_boxes only steer the cx sort -- no gradient, exactly as in the reference."""

    @staticmethod
    def forward(ctx, pred_logits, pred_boxes, targets_i32, lens_i32, eps, zero_infinity):
        import ctypes
        L.require_cuda(pred_logits, pred_boxes, targets_i32, lens_i32)
        B, Q, C = pred_logits.shape
        logits = pred_logits.float()
        if not (logits.stride(2) == 1 and logits.stride(0) == Q * logits.stride(1)):
            logits = logits.contiguous()
        boxes = pred_boxes.detach().float().contiguous()
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
        L.LAUNCHES += 4                    # row sums, sort, lp table, lattice, gradient: 5 kernels behind one C call
        per = nll / lens_i32.clamp(min=1).float()
        if zero_infinity:
            per = torch.where(torch.isinf(nll), torch.zeros_like(per), per)
        ctx.save_for_backward(grad)
        ctx.in_dtype = pred_logits.dtype
        return per.mean()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g).to(ctx.in_dtype), None, None, None, None, None


def ctc_loss(pred_logits, pred_boxes, targets_i32, lens_i32, eps=0.003, zero_infinity=True):
    """pred_logits (B,Q,C), pred_boxes (B,Q,4); targets_i32 (B,Lmax) class ids 0..C-1 (padding ignored), lens_i32 (B) -> scalar loss
    = nn.CTCLoss(blank=0, reduction='mean', zero_infinity)(log of the interleaved CTC view) of reference dino.py:505-544."""
    return _FusedCTCLoss.apply(pred_logits, pred_boxes, targets_i32.contiguous(), lens_i32.contiguous(), float(eps), bool(zero_infinity))


def gemm_ln(a, w, bias, residual, gamma, beta, add2=None, eps=1e-5, out=None, out2=None):
    """bf16 only, N = 256: y = LN(a @ w.T + bias (+ residual)); optional y2 = y + add2.  One tcgen05 kernel."""
    import ctypes
    M, K = a.shape
    assert w.shape[0] == 256 and a.dtype in HALF and w.dtype == a.dtype
    L.set_flavor(a.dtype)
    y = torch.empty((M, 256), dtype=a.dtype, device=a.device) if out is None else out
    y2 = None
    if add2 is not None:
        y2 = torch.empty((M, 256), dtype=a.dtype, device=a.device) if out2 is None else out2
    _call("dtlr_gemm_ln", _p(a), a.stride(0), _p(w), w.stride(0), _p(bias), _p(residual), residual.stride(0) if residual is not None else 0,
          _p(gamma), _p(beta), ctypes.c_float(eps), _p(y), y.stride(0), _p(add2), _p(y2), add2.stride(0) if add2 is not None else 0,
          M, K, _st(a))
    return (y, y2) if add2 is not None else y


def linear_ln(a, w, bias, residual, gamma, beta, add2=None):
    """Linear (+residual) + LayerNorm [+ second output]: fused tcgen05 kernel in bf16 mode, GEMM + LN kernels otherwise."""
    # fused only when the main loop is long enough to hide the two-pass LN epilogue (measured on B200: K = 256 layers are
    # epilogue-bound and run faster as GEMM + add_layernorm256; K = 2048 (FFN linear2) gains)
    if a.dtype in HALF and w.shape[0] == 256 and a.shape[1] >= LN_FUSE_MIN_K:
        return gemm_ln(a, w, bias, residual, gamma, beta, add2)
    # K <= 256 on many rows: the weight-stationary kernel normalises in its TMA-store epilogue.  Measured (B200, M = 58368,
    # CUDA-graph timing): 23 us without a residual vs 16 + 17 us un-fused; WITH a residual the 128 KB resident weight slice
    # leaves no room to prefetch residual tiles by TMA and the LSU transposition makes it 34 us vs 33 us -> un-fused then.
    if (LN_FUSE_WS and a.dtype in HALF and w.shape[0] == 256 and a.shape[1] <= 256 and residual is None and add2 is None
            and a.shape[0] >= 2 * 148 * 128):
        return gemm_ln(a, w, bias, None, gamma, beta, None)
    x = gemm(a, w, bias, residual=residual)
    return add_layernorm(x, None, gamma, beta, add2=add2)


# fused FFN block kernel (csrc/ffn.cu).  Measured in the full step (B200, B=64, same box): with the hidden chunk fed to linear2 from
# TMEM 8.77 ms vs 8.84 ms for linear1 + linear2/LN (the first version, hidden chunk through shared memory, lost 9.71 vs 9.54)
FFN_FUSED = _os.environ.get("DTLR_FFN_FUSED", "1") != "0"
# measured (one box, full step): no chunking 8.82 ms, 37888-row chunks 8.96, 18944: 9.32, 9472: 10.01 -> off by default
FFN_CHUNK_ROWS = int(_os.environ.get("DTLR_FFN_CHUNK_ROWS", "0"))
FFN_SPLIT_TAIL = _os.environ.get("DTLR_FFN_SPLIT_TAIL", "1") != "0"


_FFN_WS = {}
_FFN_WS_RETIRED = []      # outgrown buffers stay alive: a captured CUDA graph may still hold their address


def _ffn_workspace(device, nbytes):
    """zero-initialised workspace of the stream-K FFN kernel (ready flags + partial tiles), one per (device, stream)"""
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    ws = _FFN_WS.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            _FFN_WS_RETIRED.append(ws)
        ws = torch.zeros((nbytes,), dtype=torch.uint8, device=device)
        _FFN_WS[key] = ws
    return ws


def ffn_ln(x, w1, b1, w2, b2, gamma, beta, eps=1e-5, add2=None):
    """y = LN(x + W2 relu(W1 x + b1) + b2) [, y2 = y + add2]: one tcgen05 kernel in bf16 mode when d_model = 256 and the hidden
    width is a multiple of 128 (<= 2048) -- the hidden activation never reaches HBM (opt-in, see FFN_FUSED); otherwise linear1 +
    fused linear2/LayerNorm."""
    import ctypes
    M = x.shape[0]
    hid = w1.shape[0]
    if x.dtype == torch.float32 and w1.dtype in HALF:
        # split-precision mode: linear1 writes its ReLU output straight as the split operand of linear2 (no fp32 hidden activation in HBM)
        h = gemm(x, w1, b1, relu=1, out_dtype=SPLIT if SPLIT_OUT_FUSED else None)
        y = gemm(h, w2, b2, residual=x, out_dtype=torch.float32, split3=True)
        return add_layernorm(y, None, gamma, beta, add2=add2)
    if (FFN_FUSED and x.dtype in HALF and x.shape[1] == 256 and w2.shape[0] == 256 and hid % 128 == 0 and hid <= 2048
            and x.stride(1) == 1 and x.stride(0) % 8 == 0):
        y = torch.empty((M, 256), dtype=x.dtype, device=x.device)
        L.set_flavor(x.dtype)
        if L.TIMER is not None:     # bench.py: algorithmic FLOPs of the block = the two contractions, 2*M*hid*256 each
            L.TIMER("ffn", 4.0 * M * hid * 256, x.device, True)
        # wave quantisation: with more than one round of 128-row tiles the library runs the stream-K kernel (equal
        # (tile, hidden chunk) unit ranges per SM; neighbours exchange one fp32 partial tile through the workspace).  The workspace
        # starts with the ready flags, which must be ZERO on first use and are handed back zero by every call -> one zero-initialised
        # buffer per (device, stream), reused by every FFN block of the step (they run one after the other on that stream)
        lib = L.lib()
        lib.dtlr_ffn_workspace_bytes.restype = ctypes.c_longlong
        nws = int(lib.dtlr_ffn_workspace_bytes(M, hid)) if FFN_SPLIT_TAIL else 0
        ws = _ffn_workspace(x.device, nws) if nws > 0 else None
        _call("dtlr_ffn_ln_ws", _p(x), x.stride(0), _p(w1), w1.stride(0), _p(b1), _p(w2), w2.stride(0), _p(b2), _p(gamma), _p(beta),
              ctypes.c_float(eps), _p(y), y.stride(0), M, hid, _p(ws), ctypes.c_longlong(nws), _st(x))
        if nws > 0 and int(lib.dtlr_ffn_plan(M, hid)) == 1:      # the PART-tail plan (A/B flag): two more launches
            L.LAUNCHES += 2
        if L.TIMER is not None:
            L.TIMER("ffn", 0.0, x.device, False)
        return (y, add(y, add2)) if add2 is not None else y
    # experiment (opt-in): un-fused in row chunks that keep the hidden activation inside the 126 MB L2 (one reused buffer that
    # linear2 reads back before it is evicted).  The saved HBM traffic does not pay for the extra launches, the smaller M per
    # launch and the repeated weight-slice loads
    if (FFN_CHUNK_ROWS > 0 and x.dtype in HALF and w2.shape[0] == 256 and hid >= LN_FUSE_MIN_K and M > FFN_CHUNK_ROWS):
        y = torch.empty((M, 256), dtype=x.dtype, device=x.device)
        y2 = torch.empty((M, 256), dtype=x.dtype, device=x.device) if add2 is not None else None
        hbuf = torch.empty((FFN_CHUNK_ROWS, hid), dtype=x.dtype, device=x.device)
        for r0 in range(0, M, FFN_CHUNK_ROWS):
            r1 = min(M, r0 + FFN_CHUNK_ROWS)
            h = gemm(x[r0:r1], w1, b1, relu=1, out=hbuf[:r1 - r0])
            gemm_ln(h, w2, b2, x[r0:r1], gamma, beta, add2[r0:r1] if add2 is not None else None, eps, out=y[r0:r1],
                    out2=y2[r0:r1] if add2 is not None else None)
        return (y, y2) if add2 is not None else y
    return linear_ln(gemm(x, w1, b1, relu=1), w2, b2, x, gamma, beta, add2=add2)


MLP_HEAD_FUSED = _os.environ.get("DTLR_MLP_HEAD_FUSED", "1") != "0"


def mlp_head(x, l1, l2, w3_f32, b3, ref=None):
    """3-layer box MLP (256 -> 256 -> 256 -> 4) [+ box refinement against `ref`] as ONE tcgen05 kernel (csrc/ffn.cu, HEAD variant).
    x (M,256) 16-bit; l1 / l2 = (weight (256,256) 16-bit, bias fp32); w3_f32 (4,256) fp32, b3 (4) fp32; ref (M,4) fp32 or None.
    Returns fp32 (M,4): sigmoid(mlp(x) + inverse_sigmoid(ref)), or the raw mlp(x) when ref is None."""
    L.require_cuda(x, w3_f32, b3, ref)
    M = x.shape[0]
    assert x.dtype in HALF and x.shape[1] == 256 and x.stride(1) == 1 and l1[0].shape == (256, 256) and l2[0].shape == (256, 256)
    assert w3_f32.dtype == torch.float32 and w3_f32.shape == (4, 256) and w3_f32.is_contiguous() and (ref is None or ref.is_contiguous())
    L.set_flavor(x.dtype)
    out = torch.empty((M, 4), dtype=torch.float32, device=x.device)
    _call("dtlr_mlp_head", _p(x), x.stride(0), _p(l1[0]), l1[0].stride(0), _p(l1[1]), _p(l2[0]), l2[0].stride(0), _p(l2[1]), _p(w3_f32), _p(b3),
          _p(ref), _p(out), M, _st(x))
    return out


def preprocess_u8(packed_u8, offsets_i64, hw_i32, channels, B, Hmax, Wmax, mean, std):
    """GPU input stage (csrc/input.cu): packed u8 images -> (B,3,Hmax,Wmax) fp32 normalised + zero padded, (B,Hmax,Wmax) bool mask."""
    import ctypes
    L.require_cuda(packed_u8, offsets_i64, hw_i32)
    assert packed_u8.dtype == torch.uint8 and offsets_i64.dtype == torch.int64 and hw_i32.dtype == torch.int32
    dev = packed_u8.device
    out = torch.empty((B, 3, Hmax, Wmax), dtype=torch.float32, device=dev)
    mask = torch.empty((B, Hmax, Wmax), dtype=torch.bool, device=dev)
    m3 = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s3 = (ctypes.c_float * 3)(*[float(v) for v in std])
    _call("dtlr_preprocess_u8", _p(packed_u8), _p(offsets_i64), _p(hw_i32), int(channels), _p(out), _p(mask), B, Hmax, Wmax, m3, s3,
          _st(packed_u8))
    return out, mask


def resize_u8_bilinear(packed_in, meta_i64, tables_i32, tmp_u8, out_u8, B, channels, max_h, max_oh, max_ow):
    """PIL-exact 8-bit bilinear resize of a ragged batch (csrc/input.cu); buffers are filled in place."""
    L.require_cuda(packed_in, meta_i64, tables_i32, tmp_u8, out_u8)
    assert packed_in.dtype == torch.uint8 and meta_i64.dtype == torch.int64 and tables_i32.dtype == torch.int32
    _call("dtlr_resize_u8_bilinear", _p(packed_in), _p(meta_i64), _p(tables_i32), _p(tmp_u8), _p(out_u8), B, int(channels),
          int(max_h), int(max_oh), int(max_ow), _st(packed_in))
    return out_u8

"""ORACLE -- test infrastructure, NOT product code.

CPU restatement (plain torch CPU arithmetic / pure Python) of the stages either side of the DINO forward that
SURVEY §8(f) rows 2-4 widen into; every function cites the reference file:line it follows.

Pinning: tests/golden/make_golden_io.py runs the UNMODIFIED reference code (datasets/transforms.py ToTensor / Normalize /
resize size rule, util/misc.py nested_tensor_from_tensor_list, ngram/prediction_helpers.py get_new_pred_logits, and the
metric functions of evaluation.py extracted by AST without running that script's body) and commits the outputs as
tests/golden/io.npz + io_metrics.json; tests/test_oracle_io.py checks this restatement against them.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import re

import torch


def to_tensor_normalize(img_u8_hwc, mean, std):
    """torchvision F.to_tensor (datasets/transforms.py:247-249) + F.normalize (:552-558) on one (H,W,C) u8 image; a 1-channel
    image is first replicated to RGB as PIL `convert("RGB")` does (datasets/IAM.py:86-88)."""
    t = torch.as_tensor(img_u8_hwc)
    if t.ndim == 2:
        t = t[:, :, None]
    if t.shape[2] == 1:
        t = t.expand(-1, -1, 3)
    x = t.permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    m = torch.as_tensor(mean, dtype=torch.float32)[:, None, None]
    s = torch.as_tensor(std, dtype=torch.float32)[:, None, None]
    return x.sub(m).div(s)


def nested_batch(tensors, pad_to_w=None, pad_to_h=None):
    """util/misc.py:375-397: zero-pad (C,h,w) tensors to the batch max (or a larger given size), mask True on padding."""
    h = max(int(t.shape[1]) for t in tensors)
    w = max(int(t.shape[2]) for t in tensors)
    if pad_to_w is not None:
        w = max(w, pad_to_w)
    if pad_to_h is not None:
        h = max(h, pad_to_h)
    out = torch.zeros((len(tensors), tensors[0].shape[0], h, w), dtype=tensors[0].dtype)
    mask = torch.ones((len(tensors), h, w), dtype=torch.bool)
    for t, o, m in zip(tensors, out, mask):
        o[:, :t.shape[1], :t.shape[2]].copy_(t)
        m[:t.shape[1], :t.shape[2]] = False
    return out, mask


def resized_size(image_size_wh, size, max_size=None):
    """datasets/transforms.py:81-99 (get_size_with_aspect_ratio): (w,h) -> (oh, ow)."""
    w, h = image_size_wh
    if max_size is not None:
        mn, mx = float(min(w, h)), float(max(w, h))
        if mx / mn * size > max_size:
            size = int(round(max_size * mn / mx))
    if (w <= h and w == size) or (h <= w and h == size):
        return (h, w)
    if w < h:
        return (int(size * h / w), size)
    return (size, int(size * w / h))


def new_pred_logits(pred_logits, pred_boxes, multiply_pred_logits_by=1, eps=0.003):
    """ngram/prediction_helpers.py:5-46 (== models/dino/dino.py:472-502 when the multiplier is 1)."""
    _, idx = torch.sort(pred_boxes[:, :, 0])
    p = torch.gather(pred_logits, 1, idx.unsqueeze(-1).expand(-1, -1, pred_logits.shape[-1])).sigmoid() * multiply_pred_logits_by
    s = p.sum(-1)
    low = s < 1 - eps
    out = torch.zeros(p.shape[0], p.shape[1], p.shape[2] + 1)
    out[:, :, 1:] = torch.where(low[..., None], p, (1 - eps) * p / s[..., None])
    out[:, :, 0] = torch.where(low, 1 - s, torch.full_like(s, eps))
    return out


def edit_distance(a, b):
    """evaluation.py:309-327 / :371-390 (two-row Levenshtein)"""
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a):
        cur = [i + 1]
        for j, y in enumerate(b):
            cur.append(min(prev[j + 1] + 1, cur[j] + 1, prev[j] + (x != y)))
        prev = cur
    return prev[-1]


def cer(pred, gt):
    """evaluation.py:296-334"""
    return edit_distance(pred, gt) / max(len(gt), 1)


def wer(pred_words, gt_words):
    """evaluation.py:358-397"""
    return edit_distance(pred_words, gt_words) / max(len(gt_words), 1)


def split_words(labels, charset):
    """evaluation.py:400-411"""
    sp = charset.index(" ")
    words, word = [], []
    for v in labels:
        if v == sp:
            if word:
                words.append(word)
                word = []
        else:
            word.append(v)
    if word:
        words.append(word)
    return words


def clean_string(s):
    """evaluation.py:430-450 (process_pred_string)"""
    for a, b in (("B B C", "BBC"), ("I T V", "ITV"), ("  ", " "), (" -", "-"), ("- ", "-"), (" .", "."), (" ,", ",")):
        s = s.replace(a, b)
    s = re.sub(r"(\d), (\d)", r"\1,\2", s)
    s = s.replace(" '", "'").replace("' ", "'")
    s = re.sub(r"(?<=\S)€(?=\S)", " € ", s)
    s = re.sub(r"(?<!\.)\.\.(?!\.)", ".", s)
    return s.replace(",,", ",")


# ---------------------------------------------------------------------------------------------------------------------
# PIL's 8-bit bilinear resize (what torchvision F.resize does to the PIL images of datasets/transforms.py:107-108 ->
# Image.resize(size, BILINEAR)).  Third-party arithmetic: Pillow (12.2 installed here; `requirements.txt` pins nothing),
# src/libImaging/Resample.c -- precompute_coeffs / normalize_coeffs_8bpc / ImagingResampleHorizontal_8bpc / ...Vertical_8bpc:
# a triangle filter whose support grows with the down-scale factor (antialiasing), coefficients normalised in double precision and
# rounded to 22-bit fixed point, a horizontal pass into an 8-bit intermediate, then a vertical pass.  Restated here in numpy and
# pinned bit-for-bit against the installed PIL by tests/test_oracle_io.py.
PRECISION_BITS = 32 - 8 - 2


def resample_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs (bilinear: support 1.0) + normalize_coeffs_8bpc.
    Returns (xmin[out], count[out], coef[out, ksize] int32)."""
    import math
    import numpy as np
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    xmin_a = np.zeros(out_size, np.int32)
    cnt_a = np.zeros(out_size, np.int32)
    coef = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)          # C (int) cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = []
        for x in range(xmax):
            t = (x + xmin - center + 0.5) * ss
            t = -t if t < 0.0 else t
            w.append(1.0 - t if t < 1.0 else 0.0)
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            coef[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        xmin_a[xx], cnt_a[xx] = xmin, xmax
    return xmin_a, cnt_a, coef


def _resample_axis(img, out_size, axis):
    """one 8bpc pass along `axis` of an (H,W,C) u8 array"""
    import numpy as np
    in_size = img.shape[axis]
    xmin, cnt, coef = resample_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        ss = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for k in range(cnt[xx]):
            ss += src[xmin[xx] + k] * int(coef[xx, k])
        out[xx] = np.clip(ss >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bilinear_u8(img, out_h, out_w):
    """Image.resize((out_w, out_h), BILINEAR) on an (H,W) or (H,W,C) u8 array: horizontal pass first (only if the width
    changes), then the vertical pass (only if the height changes) -- Resample.c ImagingResampleInner."""
    import numpy as np
    a = np.asarray(img)
    squeeze = a.ndim == 2
    if squeeze:
        a = a[:, :, None]
    if out_w != a.shape[1]:
        a = _resample_axis(a, out_w, 1)
    if out_h != a.shape[0]:
        a = _resample_axis(a, out_h, 0)
    return a[:, :, 0] if squeeze else a


def nms_decode(pred_logits, pred_boxes, th, nm, num_select=900):
    """evaluation.py:94-115 (args.NMS_inference) on top of models/dino/dino.py:1000-1034 (PostProcess with target size (1,1)):
    top-`num_select` (query, class) pairs by sigmoid score, xyxy boxes, greedy class-agnostic NMS at IoU `nm` (torchvision.ops.nms:
    suppress when IoU > threshold), keep scores > th, order by box centre x.  One image (B = 1 like the reference's loop).
    Returns the class ids in reading order."""
    prob = pred_logits[0].sigmoid().reshape(-1)
    C = pred_logits.shape[2]
    scores, idx = torch.topk(prob, min(num_select, prob.numel()))
    q, labels = idx // C, idx % C
    cx, cy, w, h = pred_boxes[0][q].unbind(-1)
    boxes = torch.stack((cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h), -1)
    keep = []
    order = torch.argsort(scores, descending=True, stable=True).tolist()
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    alive = [True] * len(order)
    for a_i, a in enumerate(order):
        if not alive[a_i]:
            continue
        keep.append(a)
        for b_i in range(a_i + 1, len(order)):
            if alive[b_i]:
                b = order[b_i]
                iw = (torch.min(boxes[a, 2], boxes[b, 2]) - torch.max(boxes[a, 0], boxes[b, 0])).clamp(min=0)
                ih = (torch.min(boxes[a, 3], boxes[b, 3]) - torch.max(boxes[a, 1], boxes[b, 1])).clamp(min=0)
                inter = iw * ih
                if inter / (area[a] + area[b] - inter) > nm:
                    alive[b_i] = False
    keep = torch.tensor(keep, dtype=torch.long)
    s, l, bx = scores[keep], labels[keep], boxes[keep]
    sel = s > th
    centre = (bx[sel][:, 0] + bx[sel][:, 2]) / 2
    return l[sel][torch.sort(centre)[1]].tolist()

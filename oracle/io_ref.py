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

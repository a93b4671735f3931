"""Bucketed, batched evaluation over real-resolution line images.

The reference's evaluation loop (evaluation.py:486-535) runs one image at a time -- `model.cuda()(image[None].cuda())`, a
dozen host syncs per image in `convert_output_to_pred` (evaluation.py:116-160), `editdistance` on the critical path.  Here:
  * lines are sorted by resized width and cut into batches whose members share one padded width bucket
    (`bucket_batches`): little padding, few distinct shapes (one captured CUDA graph per bucket);
  * a batch crosses PCIe as packed u8 and is normalised / padded on the GPU (`input.GpuPreprocessor`);
  * the decode is the fused kernel pair of csrc/decode.cu with evaluation.py's eps = 0.03 / C (evaluation.py:141) and only
    the int32 frame labels come back; frames of batch i are downloaded while batch i+1 computes;
  * CER / WER accumulate on the host off the critical path, with the reference's own definitions: `character_error_rate`
    (evaluation.py:296-334), `word_error_rate` (:358-397), `split_labels_into_words` (:400-411), `process_pred_string`
    (:430-450) and the running "DAN CER" sum(dist) / sum(len) of evaluation.py:519-529.
"""
import re

import torch

from . import ops
from .input import GpuPreprocessor


# ------------------------------------------------------------------------------------------------ metrics (host)
def levenshtein_distance(s1, s2):
    """the reference's edit distance (evaluation.py:309-327; same value as `editdistance.eval`), on strings or lists."""
    if len(s1) < len(s2):
        s1, s2 = s2, s1
    if len(s2) == 0:
        return len(s1)
    previous_row = list(range(len(s2) + 1))
    for i, c1 in enumerate(s1):
        current_row = [i + 1]
        for j, c2 in enumerate(s2):
            current_row.append(min(previous_row[j + 1] + 1, current_row[j] + 1, previous_row[j] + (c1 != c2)))
        previous_row = current_row
    return previous_row[-1]


def character_error_rate(predicted_str, gt_str):
    """evaluation.py:296-334"""
    return levenshtein_distance(predicted_str, gt_str) / max(len(gt_str), 1)


def word_error_rate(predicted_words, gt_words):
    """evaluation.py:358-397"""
    return levenshtein_distance(predicted_words, gt_words) / max(len(gt_words), 1)


def split_labels_into_words(labels, charset):
    """evaluation.py:400-411: label ids -> list of words (lists of ids), split at the charset's space."""
    space = charset.index(" ")
    words, word = [], []
    for label in labels:
        if label == space:
            if word:
                words.append(word)
                word = []
        else:
            word.append(label)
    if word:
        words.append(word)
    return words


def process_pred_string(s):
    """evaluation.py:430-450 (the reference applies it to BOTH strings, :521-522)"""
    s = s.replace("B B C", "BBC")
    s = s.replace("I T V", "ITV")
    s = s.replace("  ", " ")
    s = s.replace(" -", "-")
    s = s.replace("- ", "-")
    s = s.replace(" .", ".")
    s = s.replace(" ,", ",")
    s = re.sub(r"(\d), (\d)", r"\1,\2", s)
    s = s.replace(""" '""", "'")
    s = s.replace("""' """, "'")
    s = re.sub(r"(?<=\S)€(?=\S)", " € ", s)
    s = re.sub(r"(?<!\.)\.\.(?!\.)", ".", s)
    s = s.replace(",,", ",")
    return s


# ------------------------------------------------------------------------------------------------ batching (host)
def bucket_batches(widths, batch_size, width_multiple=32, max_pad_frac=None):
    """indices sorted by width, cut into batches of <= batch_size whose widths round up to the same multiple of
    `width_multiple` (so a batch pads by < width_multiple columns beyond its widest line and every batch shape is one of a few
    buckets).  With `max_pad_frac` a batch is also cut when its narrowest line would be padded by more than that fraction.
    Returns a list of index lists; every index appears exactly once."""
    if batch_size < 1 or width_multiple < 1:
        raise ValueError("batch_size and width_multiple must be >= 1")
    order = sorted(range(len(widths)), key=lambda i: (widths[i], i))
    batches, cur, cur_bucket = [], [], None
    for i in order:
        bucket = (int(widths[i]) + width_multiple - 1) // width_multiple
        cut = cur and (len(cur) >= batch_size or bucket != cur_bucket)
        if cur and not cut and max_pad_frac is not None:
            cut = (bucket * width_multiple - widths[cur[0]]) > max_pad_frac * bucket * width_multiple
        if cut:
            batches.append(cur)
            cur = []
        if not cur:
            cur_bucket = bucket
        cur.append(i)
    if cur:
        batches.append(cur)
    return batches


def frames_to_labels(frames_row):
    """one row of int32 frame labels (0 = blank, c+1 = class c) -> class ids (engine.py:523-529 / evaluation.py:152-158)"""
    return [int(v) - 1 for v in frames_row if int(v) != 0]


# ------------------------------------------------------------------------------------------------ the loop
class LineEvaluator:
    """model: dtlr_b200 DINO in eval mode on a CUDA device; charset: list of characters (class c -> charset[c])."""

    def __init__(self, model, charset, batch_size=64, width_multiple=32, eps=None):
        self.model = model
        self.charset = list(charset)
        self.batch_size = batch_size
        self.width_multiple = width_multiple
        self.eps = eps                      # None -> 0.03 / num_classes like evaluation.py:141
        self.device = next(model.parameters()).device
        self.prep = GpuPreprocessor(self.device, pad_w_multiple=width_multiple)
        self.down = torch.cuda.Stream(device=self.device)
        self._host = {}                     # pinned result buffers: (shape, parity) -> tensor (2-deep ring per shape)

    @torch.no_grad()
    def predict(self, images):
        """images: list of resized u8 (H,W[,1|3]) arrays.  Returns the predicted class-id list of every image, input order."""
        widths = [int(im.shape[1]) for im in images]
        batches = bucket_batches(widths, self.batch_size, self.width_multiple)
        preds = [None] * len(images)
        compute = torch.cuda.current_stream(self.device)
        pending = None                      # (indices, pinned frames, event) of the previous batch
        for n, idx in enumerate(batches):
            samples = self.prep([images[i] for i in idx])
            out = self.model(samples)
            C = out["pred_logits"].shape[-1]
            frames = ops.ctc_decode(out["pred_logits"], out["pred_boxes"], self.eps if self.eps is not None else 0.03 / C)
            key = (tuple(frames.shape), n & 1)
            if key not in self._host:
                self._host[key] = torch.empty(frames.shape, dtype=frames.dtype).pin_memory()
            host = self._host[key]          # its previous user (batch n-2) was collected before batch n-1 was enqueued
            self.down.wait_stream(compute)
            with torch.cuda.stream(self.down):
                host.copy_(frames, non_blocking=True)
                frames.record_stream(self.down)
                ev = torch.cuda.Event()
                ev.record(self.down)
            if pending is not None:
                self._collect(pending, preds)
            pending = (idx, host, ev)
        if pending is not None:
            self._collect(pending, preds)
        return preds

    @staticmethod
    def _collect(pending, preds):
        idx, host, ev = pending
        ev.synchronize()
        arr = host.numpy()
        for r, i in enumerate(idx):             # vectorised blank removal: only the kept labels become Python ints
            row = arr[r]
            preds[i] = (row[row != 0] - 1).tolist()

    def evaluate(self, images, gt_labels):
        """gt_labels: list of class-id lists.  Returns dict(cer=sum(dist)/sum(len) over post-processed strings (the "DAN CER" of
        evaluation.py:529), cer_txt=mean per-line raw CER (:519-520), wer=mean per-line WER (:531-535), preds, pred_strs)."""
        preds = self.predict(images)
        cs = self.charset
        dist_sum = len_sum = 0
        cer_txt, wer = [], []
        pred_strs = []
        for p, g in zip(preds, gt_labels):
            ps = "".join(cs[c] for c in p)
            gs = "".join(cs[int(c)] for c in g)
            pred_strs.append(ps)
            cer_txt.append(levenshtein_distance(gs, ps) / max(len(gs), 1))
            pg, pp = process_pred_string(gs), process_pred_string(ps)
            dist_sum += levenshtein_distance(pg, pp)
            len_sum += len(pg)
            if " " in cs:
                # argument order as the reference calls it (evaluation.py:533): word_error_rate(gt_split, pred_split)
                wer.append(word_error_rate(split_labels_into_words([int(c) for c in g], cs), split_labels_into_words(p, cs)))
        n = max(len(preds), 1)
        return {"cer": dist_sum / max(len_sum, 1), "cer_txt": sum(cer_txt) / n, "wer": sum(wer) / max(len(wer), 1),
                "preds": preds, "pred_strs": pred_strs}

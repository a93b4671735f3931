"""Bucketed, batched evaluation over real-resolution line images.

The reference's evaluation loop (evaluation.py:486-535) runs one image at a time -- `model.cuda()(image[None].cuda())`, a
dozen host syncs per image in `convert_output_to_pred` (evaluation.py:116-160), `editdistance` on the critical path.  Here:
  * lines are cut into batches whose members share one padded (width, height) bucket (`bucket_batches`; the reference's
    evaluation resize caps the long side at 1333, so real lines share the width and differ in height): little padding, few
    distinct shapes (one captured CUDA graph per bucket);
  * a batch crosses PCIe as packed u8 and is normalised / padded on the GPU (`input.GpuPreprocessor`) on a side stream, under
    the forward of the previous batch;
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
def bucket_batches(widths, batch_size, width_multiple=32, max_pad_frac=None, heights=None, height_multiple=8):
    """indices cut into batches of <= batch_size whose members share one padded-size bucket: widths that round up to the same
    multiple of `width_multiple` and -- when `heights` is given (the reference's evaluation resize caps the long side at 1333, so
    real lines share the width and differ in height) -- heights that round up to the same multiple of `height_multiple`.  Within
    a bucket lines are taken in ascending (height, width) order, so a batch pads by less than one multiple beyond its largest line
    and every batch shape is one of a few buckets.  With `max_pad_frac` a batch is also cut when its narrowest line would be
    padded by more than that fraction of the bucket width.  Returns a list of index lists; every index appears exactly once."""
    if batch_size < 1 or width_multiple < 1 or height_multiple < 1:
        raise ValueError("batch_size and the size multiples must be >= 1")
    n = len(widths)
    if heights is not None and len(heights) != n:
        raise ValueError("heights and widths differ in length")

    def bucket(i):
        wb = (int(widths[i]) + width_multiple - 1) // width_multiple
        hb = (int(heights[i]) + height_multiple - 1) // height_multiple if heights is not None else 0
        return (wb, hb)

    order = sorted(range(n), key=lambda i: (bucket(i), int(heights[i]) if heights is not None else 0, int(widths[i]), i))
    batches, cur, cur_bucket = [], [], None
    for i in order:
        bk = bucket(i)
        cut = cur and (len(cur) >= batch_size or bk != cur_bucket)
        if cur and not cut and max_pad_frac is not None:
            cut = (bk[0] * width_multiple - min(widths[j] for j in cur)) > max_pad_frac * bk[0] * width_multiple
        if cut:
            batches.append(cur)
            cur = []
        if not cur:
            cur_bucket = bk
        cur.append(i)
    if cur:
        batches.append(cur)
    return batches


def shard_batches(batches, costs, rank, world):
    """multi-GPU evaluation (lines are independent -- no data-path collective): deterministic assignment of whole
    batches to ranks, largest cost first onto the least-loaded rank (costs: e.g. padded pixels per batch), so every rank computes
    the same assignment locally and the per-rank loads differ by at most one batch.  Returns this rank's batches, in the original
    relative order."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world")
    if len(costs) != len(batches):
        raise ValueError("one cost per batch")
    load = [0] * world
    owner = [0] * len(batches)
    for b in sorted(range(len(batches)), key=lambda i: (-costs[i], i)):
        r = min(range(world), key=lambda k: (load[k], k))
        owner[b] = r
        load[r] += costs[b]
    return [batches[i] for i in range(len(batches)) if owner[i] == rank]


def gather_predictions(preds, group=None):
    """every rank passes its list with None at the lines it did not evaluate; returns the complete list on every rank
    (`all_gather_object` of the host-side label lists -- the only communication of a sharded evaluation)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return preds
    mine = {i: p for i, p in enumerate(preds) if p is not None}
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, mine, group=group)
    out = [None] * len(preds)
    for part in parts:
        for i, p in part.items():
            if out[i] is not None:
                raise RuntimeError("line %d was evaluated by two ranks" % i)
            out[i] = p
    missing = [i for i, p in enumerate(out) if p is None]
    if missing:
        raise RuntimeError("no rank evaluated lines %s..." % missing[:5])
    return out


def frames_to_labels(frames_row):
    """one row of int32 frame labels (0 = blank, c+1 = class c) -> class ids (engine.py:523-529 / evaluation.py:152-158)"""
    return [int(v) - 1 for v in frames_row if int(v) != 0]


def nms_decode(outputs, postprocessor, TH, NM, num_select=900):
    """the reference's alternative decode, evaluation.py:94-115 (`args.NMS_inference`): PostProcess with `num_select` (query, class)
    pairs and class-agnostic NMS at IoU `NM` on unit-size images, keep scores > `TH`, read the labels in order of box centre x.
    Batched (the reference handles image 0 only); returns one class-id list per image.  Mutates the post-processor's
    `num_select` / `nms_iou_threshold` exactly as the reference does (evaluation.py:97-98)."""
    logits = outputs["pred_logits"]
    postprocessor.num_select = min(num_select, logits.shape[1] * logits.shape[2])
    postprocessor.nms_iou_threshold = NM
    sizes = torch.ones((logits.shape[0], 2), dtype=torch.float32, device=logits.device)
    if getattr(postprocessor, "fused", False) and logits.is_cuda and postprocessor.num_select <= 1024:
        # two kernels per batch (csrc/select.cu: radix-select top-k + box conversion; NMS bit matrix + greedy scan + threshold + cx order)
        # and ONE device->host copy, instead of topk / gather / per-image NMS / boolean indexing / sort with a host sync each
        labels, counts = postprocessor.read(outputs, sizes, TH)
        labels, counts = labels.cpu(), counts.cpu().tolist()
        return [labels[i, :n].tolist() for i, n in enumerate(counts)]
    res = postprocessor(outputs, sizes)
    preds = []
    for r in res:
        x0, _, x1, _ = r["boxes"].unbind(-1)
        sel = r["scores"] > TH
        order = torch.sort(((x0 + x1) / 2)[sel])[1]
        preds.append(r["labels"].long()[sel][order].tolist())
    return preds


# ------------------------------------------------------------------------------------------------ the loop
class LineEvaluator:
    """model: dtlr_b200 DINO in eval mode on a CUDA device; charset: list of characters (class c -> charset[c])."""

    def __init__(self, model, charset, batch_size=64, width_multiple=32, height_multiple=8, eps=None, nms=None):
        """nms=(postprocessor, TH, NM) switches the decode to the reference's `args.NMS_inference` branch (evaluation.py:94-115)."""
        self.model = model
        self.nms = nms
        self.charset = list(charset)
        self.batch_size = batch_size
        self.width_multiple = width_multiple
        self.height_multiple = height_multiple
        self.eps = eps                      # None -> 0.03 / num_classes like evaluation.py:141
        self.device = next(model.parameters()).device
        # one captured CUDA graph per (batch size, width bucket): keep them all alive over a data set
        model.max_cuda_graphs = max(getattr(model, "max_cuda_graphs", 8), 48)
        self.prep = GpuPreprocessor(self.device, pad_w_multiple=width_multiple, pad_h_multiple=height_multiple)
        self.up = torch.cuda.Stream(device=self.device)         # H2D + input kernel of batch i+1 overlap the forward of batch i
        self.down = torch.cuda.Stream(device=self.device)
        self._host = {}                     # pinned result buffers: (shape, parity) -> tensor (2-deep ring per shape)

    def batches(self, images):
        """the index lists `predict` runs, in order"""
        return bucket_batches([int(im.shape[1]) for im in images], self.batch_size, self.width_multiple,
                              heights=[int(im.shape[0]) for im in images], height_multiple=self.height_multiple)

    @torch.no_grad()
    def predict(self, images, rank=0, world=1):
        """images: list of resized u8 (H,W[,1|3]) arrays.  Returns the predicted class-id list of every image, input order.
        With world > 1 only this rank's share of the batches is evaluated (None elsewhere; `gather_predictions` completes it)."""
        batches = self.batches(images)
        if world > 1:
            wm, hm = self.width_multiple, self.height_multiple
            costs = [len(b) * (-(-max(int(images[i].shape[0]) for i in b) // hm) * hm) * (-(-max(int(images[i].shape[1]) for i in b) // wm) * wm)
                     for b in batches]
            batches = shard_batches(batches, costs, rank, world)
        preds = [None] * len(images)
        compute = torch.cuda.current_stream(self.device)
        pending = None                      # (indices, pinned frames, event) of the previous batch
        for n, idx in enumerate(batches):
            samples = self.prep([images[i] for i in idx], stream=self.up)
            out = self.model(samples)
            if self.nms is not None:            # alternative decode: torch ops + host lists (synchronises per batch)
                for i, p in zip(idx, nms_decode(out, *self.nms)):
                    preds[i] = p
                continue
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

    def evaluate(self, images, gt_labels, distributed=False):
        """gt_labels: list of class-id lists.  Returns dict(cer=sum(dist)/sum(len) over post-processed strings (the "DAN CER" of
        evaluation.py:529), cer_txt=mean per-line raw CER (:519-520), wer=mean per-line WER (:531-535), preds, pred_strs).
        distributed=True (torch.distributed initialised, one process per GPU, every rank passes the SAME lists): the batches are
        sharded over the ranks, the label lists gathered on the host, and every rank returns the full metrics."""
        if distributed:
            import torch.distributed as dist
            preds = gather_predictions(self.predict(images, dist.get_rank(), dist.get_world_size()))
        else:
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

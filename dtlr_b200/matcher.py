"""Hungarian matcher of the DINO detection loss: mirror of reference models/dino/matcher.py:25-96 with the
assignment solved on the GPU (dtlr_lsap) instead of scipy on a CPU copy of the cost matrix.  Also the box utilities the
losses need (reference util/box_ops.py:9-64), restated without torchvision."""
import ctypes

import torch
from torch import nn

from . import _lib as L


def box_cxcywh_to_xyxy(x):
    x_c, y_c, w, h = x.unbind(-1)
    return torch.stack([(x_c - 0.5 * w), (y_c - 0.5 * h), (x_c + 0.5 * w), (y_c + 0.5 * h)], dim=-1)


def box_xyxy_to_cxcywh(x):
    x0, y0, x1, y1 = x.unbind(-1)
    return torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, (x1 - x0), (y1 - y0)], dim=-1)


def _area(b):
    return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])


def box_iou(boxes1, boxes2):
    """reference util/box_ops.py:24-38 (union + 1e-6)"""
    area1, area2 = _area(boxes1), _area(boxes2)
    lt = torch.max(boxes1[:, None, :2], boxes2[:, :2])
    rb = torch.min(boxes1[:, None, 2:], boxes2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    union = area1[:, None] + area2 - inter
    return inter / (union + 1e-6), union


def generalized_box_iou(boxes1, boxes2):
    """reference util/box_ops.py:41-64"""
    assert (boxes1[:, 2:] >= boxes1[:, :2]).all()
    assert (boxes2[:, 2:] >= boxes2[:, :2]).all()
    iou, union = box_iou(boxes1, boxes2)
    lt = torch.min(boxes1[:, None, :2], boxes2[:, :2])
    rb = torch.max(boxes1[:, None, 2:], boxes2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    area = wh[:, :, 0] * wh[:, :, 1]
    return iou - (area - union) / (area + 1e-6)


def _lsap_target_major(cost_tm, t_cnt, B, Tmax):
    """cost_tm fp32 CUDA (P, Tmax, Q) -> int32 (P, Tmax) query of each target.  C ABI: dtlr_lsap."""
    P, _, Q = cost_tm.shape
    out = torch.empty((P, Tmax), dtype=torch.int32, device=cost_tm.device)
    with torch.cuda.device(cost_tm.device):
        rc = L.lib().dtlr_lsap(L.ptr(cost_tm), P, B, Q, L.ptr(t_cnt), Tmax, L.ptr(out), L.stream_ptr(cost_tm.device))
    L.check(rc, "dtlr_lsap")
    return out


def lsap_gpu(cost, sizes):
    """scipy.optimize.linear_sum_assignment for a batch: cost fp32 CUDA (B, Q, sum(sizes)) in the reference's layout (image b
    owns the column block of its own targets).  Returns for every image the query assigned to each of its targets (int64
    CUDA tensor of length sizes[b])."""
    L.require_cuda(cost)
    B, Q, Ttot = cost.shape
    Tmax = max(sizes) if sizes else 0
    if Tmax == 0:
        return [torch.empty(0, dtype=torch.int64, device=cost.device) for _ in sizes]
    cost_tm = torch.zeros((B, Tmax, Q), dtype=torch.float32, device=cost.device)
    o = 0
    for b, n in enumerate(sizes):
        cost_tm[b, :n] = cost[b, :, o:o + n].float().t()
        o += n
    t_cnt = torch.tensor(list(sizes), dtype=torch.int32, device=cost.device)
    out = _lsap_target_major(cost_tm, t_cnt, B, Tmax)
    return [out[b, :n].long() for b, n in enumerate(sizes)]


class HungarianMatcher(nn.Module):
    """reference matcher.py:25-96.  forward returns, per image, (query indices ascending, matching target indices) exactly like
    scipy.optimize.linear_sum_assignment on the (Q x T) cost block.  `match_layers` solves every decoder layer of a training
    step in one pair of launches (cost + assignment for layers x B problems); SetCriterion uses it when present."""

    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1, focal_alpha=0.25):
        super().__init__()
        self.cost_class, self.cost_bbox, self.cost_giou = cost_class, cost_bbox, cost_giou
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"
        self.focal_alpha = focal_alpha
        self.cpu_indices = True        # the reference hands CPU int64 index tensors to the losses; False keeps them on the GPU

    @torch.no_grad()
    def cost_matrix(self, outputs, targets):
        """the reference's full (bs, Q, sum T) cost tensor (matcher.py:57-88), torch ops; kept for inspection and tests --
        forward() computes only the per-image blocks with dtlr_match_cost."""
        bs, num_queries = outputs["pred_logits"].shape[:2]
        out_prob = outputs["pred_logits"].flatten(0, 1).float().sigmoid()
        out_bbox = outputs["pred_boxes"].flatten(0, 1).float()
        tgt_ids = torch.cat([v["labels"] for v in targets])
        tgt_bbox = torch.cat([v["boxes"] for v in targets]).float()
        alpha, gamma = self.focal_alpha, 2.0
        neg_cost_class = (1 - alpha) * (out_prob ** gamma) * (-(1 - out_prob + 1e-8).log())
        pos_cost_class = alpha * ((1 - out_prob) ** gamma) * (-(out_prob + 1e-8).log())
        cost_class = pos_cost_class[:, tgt_ids] - neg_cost_class[:, tgt_ids]
        cost_bbox = torch.cdist(out_bbox, tgt_bbox, p=1)
        cost_giou = -generalized_box_iou(box_cxcywh_to_xyxy(out_bbox), box_cxcywh_to_xyxy(tgt_bbox))
        C = self.cost_bbox * cost_bbox + self.cost_class * cost_class + self.cost_giou * cost_giou
        return C.view(bs, num_queries, -1)

    @torch.no_grad()
    def block_costs(self, layers, targets):
        """layers: list of {pred_logits (B,Q,C), pred_boxes (B,Q,4)} -> (cost (len(layers)*B, Tmax, Q) fp32 target-major, t_cnt,
        sizes, Tmax).  C ABI: dtlr_match_cost."""
        logits = torch.stack([o["pred_logits"].float() for o in layers]).contiguous()
        boxes = torch.stack([o["pred_boxes"].float() for o in layers]).contiguous()
        L.require_cuda(logits, boxes)
        nl, B, Q, C = logits.shape
        dev = logits.device
        sizes = [len(v["boxes"]) for v in targets]
        Tmax = max(sizes) if sizes else 0
        offs, o = [], 0
        for n in sizes:
            offs.append(o)
            o += n
        t_cnt = torch.tensor(sizes, dtype=torch.int32, device=dev)
        if Tmax == 0:
            return None, t_cnt, sizes, 0
        if Tmax > Q:
            raise L.DtlrError("HungarianMatcher: an image has %d targets but there are only %d queries" % (Tmax, Q))
        t_off = torch.tensor(offs, dtype=torch.int32, device=dev)
        tgt_ids = torch.cat([v["labels"] for v in targets]).to(dev, torch.int64).contiguous()
        tgt_bbox = torch.cat([v["boxes"] for v in targets]).to(dev, torch.float32).contiguous()
        cost = torch.empty((nl * B, Tmax, Q), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = L.lib().dtlr_match_cost(L.ptr(logits), L.ptr(boxes), L.ptr(tgt_ids), L.ptr(tgt_bbox), L.ptr(t_off), L.ptr(t_cnt),
                                         nl * B, B, Q, C, Tmax, ctypes.c_float(self.cost_class), ctypes.c_float(self.cost_bbox),
                                         ctypes.c_float(self.cost_giou), ctypes.c_float(self.focal_alpha), L.ptr(cost),
                                         L.stream_ptr(dev))
        L.check(rc, "dtlr_match_cost")
        return cost, t_cnt, sizes, Tmax

    @torch.no_grad()
    def match_layers(self, layers, targets):
        """one matching per entry of `layers` (same targets), two kernel launches in total"""
        B = layers[0]["pred_logits"].shape[0]
        cost, t_cnt, sizes, Tmax = self.block_costs(layers, targets)
        dev = layers[0]["pred_logits"].device
        if Tmax == 0:
            e = torch.empty(0, dtype=torch.int64, device="cpu" if self.cpu_indices else dev)
            return [[(e, e) for _ in sizes] for _ in layers]
        q_of_t = _lsap_target_major(cost, t_cnt, B, Tmax).view(len(layers), B, Tmax).long()
        # scipy returns the row (query) indices ascending with the matching column (target) indices
        key = torch.where(q_of_t >= 0, q_of_t, torch.full_like(q_of_t, 1 << 40))
        src, tgt = torch.sort(key, dim=-1)
        if self.cpu_indices:
            src, tgt = src.cpu(), tgt.cpu()
        return [[(src[l, b, :n], tgt[l, b, :n]) for b, n in enumerate(sizes)] for l in range(len(layers))]

    @torch.no_grad()
    def forward(self, outputs, targets):
        return self.match_layers([outputs], targets)[0]


def build_matcher(args):
    mt = getattr(args, "matcher_type", "HungarianMatcher")
    if mt != "HungarianMatcher":
        raise NotImplementedError("dtlr_b200 implements matcher_type='HungarianMatcher' (every shipped config uses it), got %r" % mt)
    return HungarianMatcher(cost_class=args.set_cost_class, cost_bbox=args.set_cost_bbox, cost_giou=args.set_cost_giou,
                            focal_alpha=args.focal_alpha)

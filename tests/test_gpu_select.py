"""GPU: the selection kernels of csrc/select.cu -- index work, compared index-for-index:
two-stage top-k + gathers (reference deformable_transformer.py:345-353) against torch.topk / torch.gather, PostProcess
(reference dino.py:1008-1046) against its literal torch statement, and the NMS decode of reference evaluation.py:94-115 against the
label sequences the UNMODIFIED reference produced (tests/golden/io.npz)."""
import os

import numpy as np
import pytest
import torch

from dtlr_b200 import dino, evaluation, ops

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("B,S,K", [(64, 912, 900), (3, 2676, 900), (2, 627, 100), (1, 16, 16), (5, 1500, 1)])
def test_topk_select_equals_torch_topk(B, S, K):
    g = torch.Generator(device="cuda").manual_seed(S + K)
    scores = torch.randn(B, S, device="cuda", generator=g) * 5
    idx = ops.topk_select(scores, K)
    want_v, want_i = torch.topk(scores, K, dim=1)
    assert idx.dtype == torch.int64 and torch.equal(torch.gather(scores, 1, idx), want_v)
    assert torch.equal(idx, want_i)                       # distinct random floats: the order is fully determined


def test_topk_select_ties_take_the_lowest_index_and_k_gt_s_fails_like_torch():
    scores = torch.zeros(2, 100, device="cuda")
    scores[0, 50] = 1.0
    scores[1, :] = float("-inf")
    scores[1, 7] = -3.0
    idx = ops.topk_select(scores, 10)
    assert idx[0].tolist() == [50] + list(range(9)) and idx[1].tolist() == [7] + [i for i in range(10) if i != 7][:9]
    from dtlr_b200._lib import DtlrError
    with pytest.raises(DtlrError, match="out of range"):        # reference quirk Q2: a 40x704 line has 627 < 900 tokens
        ops.topk_select(torch.zeros(1, 627, device="cuda"), 900)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_select_gather_equals_torch_gathers(dtype):
    B, S, K, d = 4, 912, 900, 256
    g = torch.Generator(device="cuda").manual_seed(1)
    delta = torch.randn(B, S, 4, device="cuda", generator=g)
    prop = torch.randn(B, S, 4, device="cuda", generator=g)
    prop[:, ::17] = float("inf")                           # invalid anchors (reference utils.py:60-61)
    mem = torch.randn(B, S, d, device="cuda", generator=g).to(dtype)
    idx = torch.stack([torch.randperm(S, device="cuda", generator=g)[:K] for _ in range(B)])
    ref, box, tgt = ops.select_gather(idx, delta, prop, mem)
    e4 = idx.unsqueeze(-1).expand(-1, -1, 4)
    assert torch.allclose(ref, torch.gather(delta + prop, 1, e4).sigmoid(), rtol=1e-6, atol=1e-7)
    assert torch.allclose(box, torch.gather(prop, 1, e4).sigmoid(), rtol=1e-6, atol=1e-7)
    assert torch.equal(tgt, torch.gather(mem, 1, idx.unsqueeze(-1).expand(-1, -1, d)))


def _pp_pair(num_select, nms):
    a, b = dino.PostProcess(num_select, nms), dino.PostProcess(num_select, nms)
    a.fused, b.fused = True, False
    return a, b


@pytest.mark.parametrize("B,Q,C,K", [(3, 900, 166, 900), (2, 300, 166, 300), (2, 50, 7356, 300), (1, 7, 5, 35)])
def test_postprocess_equals_torch_statement(B, Q, C, K):
    g = torch.Generator(device="cuda").manual_seed(Q + C)
    logits = torch.randn(B, Q, C, device="cuda", generator=g) * 3 - 4
    boxes = torch.rand(B, Q, 4, device="cuda", generator=g) * 0.5 + 0.1
    sizes = torch.tensor([[40.0, 1024.0]] * B, device="cuda")
    fused, plain = _pp_pair(K, -1)
    out = {"pred_logits": logits, "pred_boxes": boxes}
    for kw in ({}, {"not_to_xyxy": True}, {"test": True}):
        got, want = fused(out, sizes, **kw), plain(out, sizes, **kw)
        for r, w in zip(got, want):
            assert r["labels"].dtype == torch.int64
            assert torch.allclose(r["scores"], w["scores"], rtol=0, atol=2e-7)
            # the two sigmoid implementations may differ in the last bit: compare labels / boxes wherever the score order is decided
            gap = torch.minimum(torch.diff(w["scores"], prepend=w["scores"][:1] + 1).abs(), torch.diff(w["scores"], append=w["scores"][-1:] - 1).abs())
            ok = gap > 1e-6
            assert ok.float().mean() > 0.5
            assert torch.equal(r["labels"][ok], w["labels"][ok])
            assert torch.allclose(r["boxes"][ok], w["boxes"][ok], rtol=1e-6, atol=1e-6)


def test_postprocess_pitched_logits_and_saturated_scores():
    """engine logits have a padded pitch (166 -> 168); saturated sigmoids (p == 1.0f) tie and must resolve to the lowest flat index"""
    buf = torch.full((1, 40, 168), -9.0, device="cuda")
    lg = buf[..., :166]
    lg[0, 5, 10] = 30.0
    lg[0, 2, 7] = 40.0
    lg[0, 2, 3] = 25.0
    boxes = torch.rand(1, 40, 4, device="cuda")
    r = dino.PostProcess(4, -1)({"pred_logits": lg, "pred_boxes": boxes}, torch.ones(1, 2, device="cuda"))[0]
    assert r["scores"][:3].tolist() == [1.0, 1.0, 1.0]
    assert r["labels"][:3].tolist() == [3, 7, 10]           # flat indices 2*166+3 < 2*166+7 < 5*166+10


def test_nms_keep_equals_reference_greedy_nms():
    g = torch.Generator(device="cuda").manual_seed(11)
    B, Q, C, K = 3, 900, 166, 900
    logits = torch.randn(B, Q, C, device="cuda", generator=g) * 3 - 4
    cx = torch.rand(B, Q, 1, device="cuda", generator=g)
    boxes = torch.cat([cx, torch.full_like(cx, 0.5), torch.rand(B, Q, 1, device="cuda", generator=g) * 0.03 + 0.005,
                       torch.rand(B, Q, 1, device="cuda", generator=g) * 0.5 + 0.3], -1)
    sizes = torch.ones(B, 2, device="cuda")
    out = {"pred_logits": logits, "pred_boxes": boxes}
    for thr in (0.5, 0.2):
        fused, plain = _pp_pair(K, thr)
        got, want = fused(out, sizes), plain(out, sizes)
        for r, w in zip(got, want):
            assert r["scores"].shape == w["scores"].shape and 0 < r["scores"].numel() < K
            assert torch.allclose(r["scores"], w["scores"], atol=2e-7) and torch.equal(r["labels"], w["labels"])
        assert evaluation.nms_decode(out, fused, 0.3, thr) == evaluation.nms_decode(out, plain, 0.3, thr)


def test_nms_decode_equals_reference_golden_label_sequences():
    """tests/golden/io.npz nms*_labels: produced by the unmodified reference function (evaluation.py:94-115 over its own PostProcess)"""
    io = np.load(os.path.join(GOLDEN, "io.npz"))
    for i in range(3):
        lg, bx = torch.from_numpy(io["nms%d_logits" % i]).cuda(), torch.from_numpy(io["nms%d_boxes" % i]).cuda()
        th, nm = io["nms%d_th_nm" % i].tolist()
        got = evaluation.nms_decode({"pred_logits": lg, "pred_boxes": bx}, dino.PostProcess(), th, nm)
        assert got == [io["nms%d_labels" % i].tolist()]
    lg = torch.cat([torch.from_numpy(io["nms%d_logits" % i]) for i in (0, 1)]).cuda()
    bx = torch.cat([torch.from_numpy(io["nms%d_boxes" % i]) for i in (0, 1)]).cuda()
    th, nm = io["nms0_th_nm"].tolist()
    got = evaluation.nms_decode({"pred_logits": lg, "pred_boxes": bx}, dino.PostProcess(), th, nm)
    assert got[0] == io["nms0_labels"].tolist()

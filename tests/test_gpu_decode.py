"""GPU: the fused CTC-view decode (dtlr_ctc_decode) against the committed reference decode (argmax frames of
reference loss_CTC's new_pred_logits, fixture dino_A_b2) and against the torch statement on random inputs, including the
evaluation.py eps variant, the renormalised branch and a wide head."""
import numpy as np
import pytest
import torch

from dtlr_b200 import dino, ops
from gpu_common import fixture, rel

pytestmark = pytest.mark.gpu


def test_matches_reference_frames_fixture():
    fx = fixture("dino_A_b2")
    logits, boxes = torch.from_numpy(fx["pred_logits"]).cuda(), torch.from_numpy(fx["pred_boxes"]).cuda()
    frames, new = ops.ctc_decode(logits, boxes, 0.003, want_new_pred=True)
    assert (frames.cpu().numpy() == fx["ctc_argmax"]).all()
    assert rel(new[:, ::4].cpu(), fx["ctc_new_pred_s"]) < 1e-5


@pytest.mark.parametrize("B,Q,C,eps,shift", [(3, 900, 166, 0.003, -6.0), (2, 986, 166, 0.03 / 166, -6.5), (2, 300, 7356, 0.003, -10.9), (1, 17, 5, 0.003, 0.0)])
def test_matches_torch_statement(B, Q, C, eps, shift):
    g = torch.Generator(device="cuda").manual_seed(Q + C)
    logits = torch.randn(B, Q, C, device="cuda", generator=g) * 2.0 + shift
    boxes = torch.rand(B, Q, 4, device="cuda", generator=g)
    frames, new = ops.ctc_decode(logits, boxes, eps, want_new_pred=True)
    ref_new = dino.ctc_view(logits, boxes, eps)
    s = logits.sigmoid().sum(-1)
    s_sorted = torch.gather(s, 1, torch.sort(boxes[:, :, 0])[1])
    clear = (s_sorted - (1 - eps)).abs() > 1e-4          # rows whose branch (s < 1-eps) does not hinge on the summation order
    assert torch.allclose(new[clear], ref_new[clear], rtol=2e-5, atol=1e-7)
    top2 = ref_new.topk(2, dim=-1)[0]
    decided = ((top2[..., 0] - top2[..., 1]) > 1e-6) & clear
    assert (frames.long() == ref_new.argmax(-1))[decided].all()
    assert ((s < 1 - eps).any() and (s >= 1 - eps).any()) or C <= 5      # both branches of the blank synthesis exercised


@pytest.mark.parametrize("C,pitch", [(166, 168), (7355, 7356), (4, 4), (1001, 1004)])
def test_pitched_rows_take_the_vector_path_and_match_contiguous_rows(C, pitch):
    """engine layout: class rows padded to a multiple of 4 floats (166 -> 168) run the 16-byte-load kernel with a scalar tail of
    C % 4 classes; same frames / new_pred_logits as the contiguous (scalar-load) call and as the torch statement."""
    g = torch.Generator(device="cuda").manual_seed(C)
    buf = torch.randn(2, 333, pitch, device="cuda", generator=g) * 2.0 - (6.0 if C < 2000 else 10.5)
    logits = buf[:, :, :C]
    boxes = torch.rand(2, 333, 4, device="cuda", generator=g)
    from dtlr_b200 import _lib
    f_vec, n_vec = ops.ctc_decode(logits, boxes, 0.003, want_new_pred=True)      # pitched rows: the 16-byte-load row kernel
    _lib.lib().dtlr_debug_flags(32768)                                            # same pitched rows on the scalar-load kernel
    try:
        f_sc2 = ops.ctc_decode(logits, boxes, 0.003)
    finally:
        _lib.lib().dtlr_debug_flags(0)
    f_sc, n_sc = ops.ctc_decode(logits.contiguous() if C % 4 else logits.contiguous()[:, :, :C], boxes, 0.003, want_new_pred=True)
    ref_new = dino.ctc_view(logits, boxes, 0.003)
    s = torch.gather(logits.sigmoid().sum(-1), 1, torch.sort(boxes[:, :, 0])[1])
    clear = (s - (1 - 0.003)).abs() > 1e-4
    top2 = ref_new.topk(2, dim=-1)[0]
    decided = ((top2[..., 0] - top2[..., 1]) > 1e-6) & clear
    assert torch.allclose(n_vec[clear], ref_new[clear], rtol=2e-5, atol=1e-7)
    assert (f_vec.long() == ref_new.argmax(-1))[decided].all() and (f_vec == f_sc)[decided].all() and (f_vec == f_sc2)[decided].all()

"""Host logic of the detection loss on CPU against the reference golden fixture: the loss functions with the
reference's own matching, and the matcher's cost matrix (solved here with scipy as the checker; the product solver is the CUDA
kernel dtlr_lsap, covered by tests/test_gpu_criterion.py)."""
import pytest
import torch
from scipy.optimize import linear_sum_assignment

from criterion_common import CASES, load_case
from dtlr_b200.dino import SetCriterion
from dtlr_b200.matcher import HungarianMatcher, generalized_box_iou, box_cxcywh_to_xyxy, lsap_gpu
from dtlr_b200._lib import DtlrError


class ReplayMatcher:
    """returns the reference's stored matching in call order: final, aux 0..n-1, interm"""

    def __init__(self, case):
        n = case["n_aux"]
        self.order = [case["indices"][n + 1]] + case["indices"][:n] + [case["indices"][n]]
        self.k = 0

    def __call__(self, outputs, targets):
        r = self.order[self.k]
        self.k += 1
        return r


@pytest.mark.parametrize("name", CASES)
def test_losses_match_reference_given_reference_matching(name):
    c = load_case(name)
    crit = SetCriterion(c["C"], matcher=ReplayMatcher(c), weight_dict={}, focal_alpha=0.25, losses=["labels", "boxes", "cardinality"])
    crit.train(c["train"])
    losses = crit(c["outputs"], c["targets"])
    assert set(losses) == set(c["losses"])
    for k, v in c["losses"].items():
        assert float(losses[k]) == pytest.approx(v, rel=1e-5, abs=1e-6), k


@pytest.mark.parametrize("name", CASES)
def test_cost_matrix_reproduces_reference_matching(name):
    c = load_case(name)
    m = HungarianMatcher(cost_class=2.0, cost_bbox=5.0, cost_giou=2.0, focal_alpha=0.25)
    C = m.cost_matrix({k: c["outputs"][k] for k in ("pred_logits", "pred_boxes")}, c["targets"])
    assert C.shape == (c["B"], c["Q"], sum(c["sizes"]))
    final = c["indices"][c["n_aux"] + 1]
    off = 0
    for b, n in enumerate(c["sizes"]):
        i, j = linear_sum_assignment(C[b, :, off:off + n].numpy())
        assert i.tolist() == final[b][0].tolist() and j.tolist() == final[b][1].tolist()
        off += n


def test_giou_properties():
    g = torch.Generator().manual_seed(0)
    a = torch.cat([torch.rand(9, 2, generator=g), torch.rand(9, 2, generator=g) * 0.3 + 0.01], -1)
    x = box_cxcywh_to_xyxy(a)
    gi = generalized_box_iou(x, x)
    assert torch.allclose(torch.diag(gi), torch.ones(9), atol=2e-3)   # union + 1e-6 (reference box_ops.py:37) keeps small boxes < 1
    assert (gi <= 1 + 1e-6).all() and (gi >= -1 - 1e-6).all()
    assert torch.allclose(gi, gi.t(), atol=1e-6)


def test_ctc_dispatch_and_eval_flag():
    c = load_case("D1")
    crit = SetCriterion(c["C"], matcher=ReplayMatcher(c), weight_dict={}, focal_alpha=0.25, losses=["labels", "boxes", "cardinality"],
                        CTC=True)
    out = crit(c["outputs"], c["targets"])
    assert set(out) == {"loss_CTC", "loss_CTC_0", "loss_CTC_1", "loss_CTC_interm"}
    out = crit(c["outputs"], c["targets"], eval=True)
    assert "loss_giou" in out and "loss_ce_interm" in out


def test_matcher_has_no_cpu_fallback():
    with pytest.raises(DtlrError):
        lsap_gpu(torch.zeros(1, 4, 2), [2])

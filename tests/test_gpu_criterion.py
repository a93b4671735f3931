"""GPU parity of the detection-loss path: dtlr_lsap against scipy.optimize.linear_sum_assignment, the
HungarianMatcher + SetCriterion against the reference golden fixture (tests/golden/criterion.npz)."""
import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

from criterion_common import CASES, load_case

pytestmark = pytest.mark.gpu


def _solve_ref(cost, sizes):
    out, off = [], 0
    for b, n in enumerate(sizes):
        i, j = linear_sum_assignment(cost[b, :, off:off + n])
        q = np.empty(n, dtype=np.int64)
        q[j] = i
        out.append((q, cost[b, i, off + j].astype(np.float64).sum()))
        off += n
    return out


@pytest.mark.parametrize("B,Q,sizes", [(1, 8, [3]), (3, 40, [5, 0, 9]), (2, 16, [16, 1]), (4, 900, [100, 37, 1, 64]),
                                       (64, 900, None)])
def test_lsap_matches_scipy(B, Q, sizes):
    from dtlr_b200.matcher import lsap_gpu
    g = torch.Generator().manual_seed(B * 1000 + Q)
    if sizes is None:
        sizes = torch.randint(20, 101, (B,), generator=g).tolist()
    cost = torch.randn(B, Q, sum(sizes), generator=g) * 3
    got = lsap_gpu(cost.cuda(), sizes)
    ref = _solve_ref(cost.numpy(), sizes)
    off = 0
    for b, n in enumerate(sizes):
        q = got[b].cpu().numpy()
        assert q.shape == (n,)
        assert len(set(q.tolist())) == n and (q >= 0).all() and (q < Q).all()
        tot = cost[b, q, off + np.arange(n)].double().sum().item() if n else 0.0
        assert tot == pytest.approx(ref[b][1], rel=1e-9, abs=1e-9)
        assert q.tolist() == ref[b][0].tolist()           # continuous random costs: the optimum is unique
        off += n


def test_lsap_with_ties_is_optimal():
    """integer costs -> many optimal matchings; the assignment must still be a valid one of minimal total cost"""
    from dtlr_b200.matcher import lsap_gpu
    g = torch.Generator().manual_seed(5)
    sizes = [12, 30, 7]
    cost = torch.randint(0, 4, (3, 48, sum(sizes)), generator=g).float()
    got = lsap_gpu(cost.cuda(), sizes)
    ref = _solve_ref(cost.numpy(), sizes)
    off = 0
    for b, n in enumerate(sizes):
        q = got[b].cpu().numpy()
        assert len(set(q.tolist())) == n
        assert cost[b, q, off + np.arange(n)].double().sum().item() == ref[b][1]
        off += n


@pytest.mark.parametrize("name", CASES)
def test_matcher_indices_match_reference(name):
    from dtlr_b200.matcher import HungarianMatcher
    c = load_case(name, "cuda")
    m = HungarianMatcher(cost_class=2.0, cost_bbox=5.0, cost_giou=2.0, focal_alpha=0.25)
    n = c["n_aux"]
    layers = [c["outputs"]] + c["outputs"]["aux_outputs"] + [c["outputs"]["interm_outputs"]]
    want = [c["indices"][n + 1]] + c["indices"][:n] + [c["indices"][n]]
    for out, ref in zip(layers, want):
        got = m({k: out[k] for k in ("pred_logits", "pred_boxes")}, c["targets"])
        for (gi, gj), (ri, rj) in zip(got, ref):
            assert gi.dtype == torch.int64 and gi.device.type == "cpu"
            assert gi.tolist() == ri.tolist() and gj.tolist() == rj.tolist()


@pytest.mark.parametrize("name", CASES)
def test_criterion_matches_reference(name):
    from dtlr_b200.dino import SetCriterion
    from dtlr_b200.matcher import HungarianMatcher
    c = load_case(name, "cuda")
    crit = SetCriterion(c["C"], matcher=HungarianMatcher(2.0, 5.0, 2.0, 0.25), weight_dict={}, focal_alpha=0.25,
                        losses=["labels", "boxes", "cardinality"])
    crit.train(c["train"])
    losses, ind = crit(c["outputs"], c["targets"], return_indices=True)
    assert set(losses) == set(c["losses"])
    for k, v in c["losses"].items():
        assert float(losses[k]) == pytest.approx(v, rel=2e-5, abs=1e-6), k
    assert len(ind) == c["n_aux"] + 2


def test_criterion_backward_through_model_outputs():
    """the detection loss drives gradients into logits and boxes (training path of main_synthetic.py)"""
    from dtlr_b200.dino import SetCriterion
    from dtlr_b200.matcher import HungarianMatcher
    c = load_case("D2", "cuda")
    out = c["outputs"]
    out["pred_logits"].requires_grad_(True)
    out["pred_boxes"].requires_grad_(True)
    crit = SetCriterion(c["C"], matcher=HungarianMatcher(2.0, 5.0, 2.0, 0.25),
                        weight_dict={"loss_ce": 1.0, "loss_bbox": 5.0, "loss_giou": 2.0}, focal_alpha=0.25,
                        losses=["labels", "boxes", "cardinality"]).train()
    losses = crit(out, c["targets"])
    total = sum(losses[k] * w for k, w in crit.weight_dict.items())
    total.backward()
    assert torch.isfinite(out["pred_logits"].grad).all() and out["pred_logits"].grad.abs().sum() > 0
    assert torch.isfinite(out["pred_boxes"].grad).all() and out["pred_boxes"].grad.abs().sum() > 0


def test_block_costs_match_reference_cost_matrix():
    """dtlr_match_cost (block diagonal, target-major) against the torch restatement of matcher.py:57-88"""
    from dtlr_b200.matcher import HungarianMatcher
    c = load_case("D3", "cuda")
    m = HungarianMatcher(2.0, 5.0, 2.0, 0.25)
    layers = [{k: c["outputs"][k] for k in ("pred_logits", "pred_boxes")}, {k: c["outputs"]["interm_outputs"][k] for k in ("pred_logits", "pred_boxes")}]
    cost, t_cnt, sizes, Tmax = m.block_costs(layers, c["targets"])
    assert cost.shape == (2 * c["B"], Tmax, c["Q"])
    for l, out in enumerate(layers):
        full = m.cost_matrix(out, c["targets"])
        off = 0
        for b, n in enumerate(sizes):
            want = full[b, :, off:off + n].t()
            got = cost[l * c["B"] + b, :n]
            assert torch.allclose(got, want, rtol=1e-5, atol=2e-5), (l, b, (got - want).abs().max().item())
            off += n


def test_match_layers_equals_per_layer_calls_and_gpu_indices():
    from dtlr_b200.matcher import HungarianMatcher
    c = load_case("D1", "cuda")
    m = HungarianMatcher(2.0, 5.0, 2.0, 0.25)
    layers = [c["outputs"]] + c["outputs"]["aux_outputs"] + [c["outputs"]["interm_outputs"]]
    layers = [{k: o[k] for k in ("pred_logits", "pred_boxes")} for o in layers]
    batched = m.match_layers(layers, c["targets"])
    m.cpu_indices = False
    for l, o in enumerate(layers):
        single = m(o, c["targets"])
        for (bi, bj), (si, sj) in zip(batched[l], single):
            assert si.is_cuda and bi.tolist() == si.tolist() and bj.tolist() == sj.tolist()

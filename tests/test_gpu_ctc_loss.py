"""GPU: the fused CTC loss (dtlr_ctc_loss, csrc/decode.cu) against the reference's literal chain (SetCriterion.loss_CTC with
fused_ctc=False = reference models/dino/dino.py:457-551 statement for statement over torch ops + F.ctc_loss): loss value and the
gradient with respect to pred_logits, both blank branches, repeated labels, empty targets, pitched logits, C = 166 and C = 7356,
and the loss of the reference-generated training fixture."""
import pytest
import torch

from dtlr_b200 import dino, ops
from gpu_common import fixture

pytestmark = pytest.mark.gpu


def _crit():
    return dino.SetCriterion(166, None, {"loss_CTC": 1.0}, 0.25, ["labels"])


def _case(B, Q, C, lens, seed, bias=-3.0, repeat=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    logits = torch.randn(B, Q, C, generator=g) * 2.0 + bias
    # rows alternate between the two blank branches of reference dino.py:491-502: sum of sigmoids well below / well above 1
    logits[:, 0::2] -= (torch.log(torch.tensor(float(C))) + 1.5)
    logits[:, 1::2] += 1.0
    logits = logits.cuda()
    boxes = torch.rand(B, Q, 4, generator=g).cuda()
    targets = []
    for n in lens:
        lab = torch.randint(0, C, (n,), generator=g)
        if repeat and n >= 4:
            lab[1] = lab[0]
            lab[3] = lab[2]
        targets.append({"labels": lab.cuda(), "boxes": torch.zeros(n, 4).cuda()})
    return logits, boxes, targets


def _both(logits, boxes, targets):
    crit = _crit()
    res = []
    for fused in (False, True):
        crit.fused_ctc = fused
        x = logits.clone().requires_grad_(True)
        bx = boxes.clone().requires_grad_(True)
        loss = crit.loss_CTC({"pred_logits": x, "pred_boxes": bx}, targets, None, None)["loss_CTC"]
        loss.backward()
        res.append((loss.detach(), x.grad.clone(), bx.grad))
    return res


@pytest.mark.parametrize("B,Q,C,lens,repeat", [(3, 60, 20, [7, 12, 3], True), (2, 300, 166, [42, 57], False),
                                               (4, 97, 33, [0, 1, 20, 40], True), (2, 120, 7356, [30, 45], False)])
def test_fused_loss_and_gradient_match_the_reference_chain(B, Q, C, lens, repeat):
    logits, boxes, targets = _case(B, Q, C, lens, seed=B * 100 + Q, repeat=repeat)
    (l0, g0, b0), (l1, g1, b1) = _both(logits, boxes, targets)
    assert torch.isfinite(l1)
    assert abs(l1.item() - l0.item()) <= 1e-4 * abs(l0.item()) + 1e-6
    scale = g0.abs().max().item()
    err = (g1 - g0).abs().max().item()
    print("ctc loss %.6f vs %.6f; grad max %.3e, max abs diff %.3e" % (l1.item(), l0.item(), scale, err))
    assert err <= 1e-3 * scale + 1e-9
    assert b1 is None or float(b1.abs().max()) == 0.0        # boxes only steer the sort


@pytest.mark.parametrize("Q,C", [(60, 20), (300, 166), (120, 7356)])
def test_low_and_high_branch_rows_are_both_exercised(Q, C):
    logits, boxes, targets = _case(2, Q, C, [10, 15], seed=5)
    s = logits.sigmoid().sum(-1)
    lo = (s < 1 - 0.003).float().mean().item()
    assert 0.2 < lo < 0.8, lo


def test_pitched_bf16_logits_and_return_preds():
    """the engine hands out logits with a padded row pitch (166 -> 168); autocast hands out bf16"""
    logits, boxes, targets = _case(2, 80, 166, [20, 33], seed=9)
    buf = torch.zeros(2, 80, 168, device="cuda")
    buf[..., :166] = logits
    crit = _crit()
    a, new, _ = crit.loss_CTC({"pred_logits": buf[..., :166], "pred_boxes": boxes}, targets, None, None, return_preds=True)
    b = crit.loss_CTC({"pred_logits": logits, "pred_boxes": boxes}, targets, None, None)
    assert torch.equal(a["loss_CTC"], b["loss_CTC"])
    assert new.shape == (2, 80, 167) and torch.allclose(new, dino.ctc_view(logits, boxes), rtol=1e-5, atol=1e-7)
    x = logits.bfloat16().requires_grad_(True)
    crit.loss_CTC({"pred_logits": x, "pred_boxes": boxes}, targets, None, None)["loss_CTC"].backward()
    assert x.grad.dtype == torch.bfloat16 and torch.isfinite(x.grad.float()).all()


def test_reference_training_fixture_loss():
    """fixture dino_T_b2: pred_logits / pred_boxes / ctc_loss produced by the UNMODIFIED reference in training mode (quirk Q3:
    2*max_len DN queries stay in the output and flow into the loss)"""
    from dtlr_b200 import synth
    fx = fixture("dino_T_b2")
    tg = [{k: v.cuda() for k, v in t.items()} for t in synth.synth_targets(2, 166, seed=3)]
    out = {"pred_logits": torch.from_numpy(fx["pred_logits"]).cuda(), "pred_boxes": torch.from_numpy(fx["pred_boxes"]).cuda()}
    loss = _crit().loss_CTC(out, tg, None, None)["loss_CTC"]
    assert abs(loss.item() - float(fx["ctc_loss"])) < 1e-3 * float(fx["ctc_loss"])


def test_gradcheck_against_finite_differences():
    """the analytic gradient of the fused kernel vs central differences of its own loss (fp32: coarse step, a few coordinates)"""
    logits, boxes, targets = _case(1, 24, 9, [6], seed=3, bias=-1.0, repeat=True)
    tt = torch.zeros(1, 6, dtype=torch.int32, device="cuda")
    tt[0] = targets[0]["labels"].int()
    ll = torch.tensor([6], dtype=torch.int32, device="cuda")
    x = logits.clone().requires_grad_(True)
    ops.ctc_loss(x, boxes, tt, ll).backward()
    g = x.grad
    h = 1e-2
    for (q, c) in [(0, 0), (5, int(tt[0, 0])), (11, int(tt[0, 2])), (3, 4), (23, int(tt[0, 5]))]:
        xp, xm = logits.clone(), logits.clone()
        xp[0, q, c] += h
        xm[0, q, c] -= h
        fd = (ops.ctc_loss(xp, boxes, tt, ll) - ops.ctc_loss(xm, boxes, tt, ll)).item() / (2 * h)
        assert abs(fd - g[0, q, c].item()) <= 2e-2 * max(abs(fd), float(g.abs().max())), (q, c, fd, g[0, q, c].item())

"""GPU: the autograd-capable module path (torch contractions + C-ABI deformable attention) against the committed
reference vectors.  fp32, TF32 off.  Tolerance 1e-3 relative-to-max (north star), observed ~1e-5."""
import numpy as np
import pytest
import torch

from dtlr_b200 import synth
from gpu_common import build_model, fixture, near_tie_mask, rel

pytestmark = pytest.mark.gpu
TOL = 1e-3


def run_modules(model, x, targets=None, force=None):
    model.use_engine = False
    st = {}
    model.transformer.debug_stages = st
    model.transformer.debug_force_topk = force
    with torch.no_grad():
        out = model(x) if targets is None else model(x, targets)
    model.transformer.debug_stages = None
    model.transformer.debug_force_topk = None
    return out, st


def check(fx, out, st, forced):
    assert rel(st["memory"][:, ::8, ::4], fx["memory_s"]) < TOL
    assert rel(st["topk_scores"], fx["topk_scores"]) < TOL
    if not forced:
        mism = st["topk_idx"].cpu().numpy() != fx["topk_idx"]
        assert (near_tie_mask(fx)[mism]).all()
    assert rel(out["pred_logits"], fx["pred_logits"]) < TOL
    assert rel(out["pred_boxes"], fx["pred_boxes"]) < TOL
    assert rel(out["aux_outputs"][4]["pred_logits"][:, ::8, ::4], fx["aux4_logits_s"]) < TOL
    assert rel(out["aux_outputs"][0]["pred_boxes"], fx["aux0_boxes"]) < TOL
    assert rel(out["interm_outputs"]["pred_logits"][:, ::8, ::4], fx["interm_logits_s"]) < TOL
    assert rel(out["interm_outputs"]["pred_boxes"], fx["interm_boxes"]) < TOL
    assert rel(out["interm_outputs_for_matching_pre"]["pred_boxes"], fx["init_box_proposal"]) < TOL


def test_config1_single_line_100_queries():
    fx = fixture("dino_P_b1")
    model, _, _ = build_model(100)
    model.eval()
    out, st = run_modules(model, synth.synth_images(1, 40, 704, seed=1).cuda())
    assert (st["topk_idx"].cpu().numpy() == fx["topk_idx"]).all()
    check(fx, out, st, forced=False)


def test_ragged_batch():
    fx = fixture("dino_R_b3")
    model, _, _ = build_model(300)
    model.eval()
    imgs = [t.cuda() for t in synth.synth_images(3, 40, 1024, seed=2, widths=fx["widths"].tolist())]
    force = torch.from_numpy(fx["topk_idx"]).long()
    out, st = run_modules(model, imgs, force=force)
    check(fx, out, st, forced=True)
    out2, st2 = run_modules(model, imgs)
    check(fx, out2, st2, forced=False) if (st2["topk_idx"].cpu().numpy() == fx["topk_idx"]).all() else None


def test_config2_shape_900_queries_and_decode():
    fx = fixture("dino_A_b2")
    model, crit, post = build_model(900)
    model.eval()
    x = synth.synth_images(2, 40, 1024, seed=0).cuda()
    out, st = run_modules(model, x, force=torch.from_numpy(fx["topk_idx"]).long())
    check(fx, out, st, forced=True)
    targets = synth.synth_targets(2, 166, seed=0)
    losses, new, _ = crit.loss_CTC(out, targets, None, None, return_preds=True)
    assert rel(new[:, ::4], fx["ctc_new_pred_s"]) < TOL
    agree = (new.argmax(-1).cpu().numpy() == fx["ctc_argmax"]).mean()
    assert agree == 1.0, "argmax character sequence differs (%.5f agreement)" % agree
    assert abs(losses["loss_CTC"].item() - float(fx["ctc_loss"])) < 1e-3 * float(fx["ctc_loss"])
    post["bbox"].num_select = 300
    res = post["bbox"](out, torch.tensor([[40.0, 1024.0]] * 2).cuda())
    assert rel(torch.stack([r["scores"] for r in res]), fx["pp_scores"]) < TOL
    assert (torch.stack([r["labels"] for r in res]).cpu().numpy() == fx["pp_labels"]).mean() > 0.99


def test_training_mode_forward_backward_Q3():
    fx = fixture("dino_T_b2")
    model, crit, _ = build_model(300)
    model.train()
    model.use_engine = False
    tg = synth.synth_targets(2, 166, seed=3)
    tg_dev = [{k: v.cuda() for k, v in t.items()} for t in tg]
    model.transformer.debug_force_topk = torch.from_numpy(fx["topk_idx"]).long()
    out = model(synth.synth_images(2, 40, 1024, seed=3).cuda(), tg_dev)
    model.transformer.debug_force_topk = None
    assert out["pred_logits"].shape[1] == int(fx["pad_size"]) + 300
    assert rel(out["pred_logits"], fx["pred_logits"]) < TOL
    assert rel(out["pred_boxes"], fx["pred_boxes"]) < TOL
    losses = crit.loss_CTC(out, tg_dev, None, None)
    assert abs(losses["loss_CTC"].item() - float(fx["ctc_loss"])) < 1e-3 * float(fx["ctc_loss"])
    losses["loss_CTC"].backward()
    g = model.transformer.encoder.layers[0].self_attn.sampling_offsets.weight.grad
    assert g is not None and torch.isfinite(g).all() and g.abs().sum() > 0
    assert model.backbone[0].body.conv1.weight.grad is None          # conv1 + layer1 frozen (reference backbone.py:79-81)


def test_sine_embed_bf16_fast_kernel_matches_reference_formula():
    """gen_sineembed_for_position (reference models/dino/utils.py:141-167) on reference boxes x valid ratios: the bf16 throughput
    kernel (SFU sin/cos, packed stores) against the exact fp32 kernel and a plain torch statement of the formula."""
    import math
    from dtlr_b200 import ops
    B, Q, L = 3, 900, 4
    g = torch.Generator(device="cuda").manual_seed(4)
    ref = torch.rand(B * Q, 4, device="cuda", generator=g)
    vr = (0.5 + 0.5 * torch.rand(B, L, 2, device="cuda", generator=g)).contiguous()
    exact = ops.sine_embed(ref, vr, B, Q, L, torch.float32)
    fast = ops.sine_embed(ref, vr, B, Q, L, torch.bfloat16).float()
    assert (fast - exact).abs().max().item() <= 2 ** -8 + 1e-6           # bf16 rounding of values in [-1, 1]
    dim_t = 10000 ** (2 * (torch.arange(128, device="cuda") // 2) / 128.0)
    scl = torch.cat([vr[:, 0], vr[:, 0]], -1)[:, None, :]                # (B,1,4) = (vx, vy, vx, vy)
    r = ref.view(B, Q, 4) * scl
    parts = []
    for comp in (r[..., 1], r[..., 0], r[..., 2], r[..., 3]):            # y, x, w, h
        p = comp[..., None] * 2 * math.pi / dim_t
        parts.append(torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), -1).flatten(-2))
    formula = torch.cat(parts, -1).view(B * Q, 512)
    assert (exact - formula).abs().max().item() < 1e-5

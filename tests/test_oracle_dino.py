"""CPU: the oracle restatement of the DINO forward (oracle/dino_ref.py) against the committed golden vectors produced
by the UNMODIFIED reference (tests/golden/make_golden.py).  Tolerance 1e-4 relative-to-max (fp32 CPU on both sides;
observed residual ~1e-6), far inside the 1e-3 the north star allows the CUDA path."""
import json
import os

import numpy as np
import pytest
import torch

from dtlr_b200 import synth
from oracle import dino_ref

TOL = 1e-4


def rel(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    fin = torch.isfinite(a) & torch.isfinite(b)
    assert (torch.isfinite(a) == torch.isfinite(b)).all()
    return ((a - b)[fin].abs().max() / b[fin].abs().max().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def sd(golden_dir):
    shapes = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))
    return shapes


def _sd(shapes, num_queries):
    shapes = dict(shapes)
    shapes["transformer.tgt_embed.weight"] = [num_queries, 256]
    return synth.synth_state_dict(shapes, seed=0)


def check_case(fx, out, st, forced):
    assert rel(st["memory"][:, ::8, ::4], fx["memory_s"]) < TOL
    assert rel(st["topk_scores"], fx["topk_scores"]) < TOL
    assert rel(st["feats"][0][:, ::16, :, ::8], fx["feat_c3_s"]) < TOL
    assert rel(st["feats"][2][:, ::32], fx["feat_c5_s"]) < TOL
    if not forced:
        assert (st["topk_idx"].numpy() == fx["topk_idx"]).all()
    assert rel(st["hs"][0][:, ::8, ::4], fx["hs0_s"]) < TOL
    assert rel(st["refs"][1], fx["ref1"]) < TOL
    assert rel(out["pred_logits"], fx["pred_logits"]) < TOL
    assert rel(out["pred_boxes"], fx["pred_boxes"]) < TOL
    assert rel(out["aux_outputs"][4]["pred_logits"][:, ::8, ::4], fx["aux4_logits_s"]) < TOL
    assert rel(out["aux_outputs"][0]["pred_boxes"], fx["aux0_boxes"]) < TOL
    assert rel(out["interm_outputs"]["pred_logits"][:, ::8, ::4], fx["interm_logits_s"]) < TOL
    assert rel(out["interm_outputs"]["pred_boxes"], fx["interm_boxes"]) < TOL
    assert rel(out["interm_outputs_for_matching_pre"]["pred_boxes"], fx["init_box_proposal"]) < TOL


def test_config1_P_single_line_100_queries(golden_dir, sd):
    """BASELINE config 1: 1x3x40x704, num_queries=100 (the shipped 900 cannot run: S=627 < 900)."""
    fx = np.load(os.path.join(golden_dir, "dino_P_b1.npz"))
    st = {}
    cfg = dino_ref.default_cfg(num_queries=100)
    out = dino_ref.dino_forward(_sd(sd, 100), cfg, synth.synth_images(1, 40, 704, seed=1), stages=st)
    check_case(fx, out, st, forced=False)


def test_ragged_batch_masks(golden_dir, sd):
    fx = np.load(os.path.join(golden_dir, "dino_R_b3.npz"))
    st = {}
    cfg = dino_ref.default_cfg(num_queries=300)
    imgs = synth.synth_images(3, 40, 1024, seed=2, widths=fx["widths"].tolist())
    out = dino_ref.dino_forward(_sd(sd, 300), cfg, imgs, stages=st)
    check_case(fx, out, st, forced=False)


def test_config2_A_900_queries_and_decode_tail(golden_dir, sd):
    """BASELINE config 2 shape (40x1024, Q=900).  900 of 912 tokens are selected and a few two-stage scores are closer
    than fp32 round-off, so the ranking is compared where it is numerically decidable and the downstream tensors with
    the reference ranking forced (parity protocol)."""
    fx = np.load(os.path.join(golden_dir, "dino_A_b2.npz"))
    cfg = dino_ref.default_cfg(num_queries=900)
    w = _sd(sd, 900)
    x = synth.synth_images(2, 40, 1024, seed=0)
    st = {}
    dino_ref.dino_forward(w, cfg, x, stages=st)
    assert rel(st["topk_scores"], fx["topk_scores"]) < TOL
    mism = st["topk_idx"].numpy() != fx["topk_idx"]
    ref_scores = np.take_along_axis(fx["topk_scores"], fx["topk_idx"].astype(np.int64), 1)
    gap = np.minimum(np.abs(np.diff(ref_scores, axis=1, prepend=np.inf)), np.abs(np.diff(ref_scores, axis=1, append=-np.inf)))
    assert (gap[mism] < 1e-4).all()            # ranking differs only inside near-ties
    st = {}
    out = dino_ref.dino_forward(w, cfg, x, stages=st, force_topk=torch.from_numpy(fx["topk_idx"]).long())
    check_case(fx, out, st, forced=True)
    # decode tail (reference dino.py:457-551, engine.py:512-530, dino.py:994-1046)
    logits, boxes = torch.from_numpy(fx["pred_logits"]), torch.from_numpy(fx["pred_boxes"])
    targets = synth.synth_targets(2, 166, seed=0)
    loss, new = dino_ref.loss_ctc(logits, boxes, targets)
    assert rel(new[:, ::4], fx["ctc_new_pred_s"]) < 1e-6
    assert (new.argmax(-1).numpy() == fx["ctc_argmax"]).all()
    assert abs(loss.item() - float(fx["ctc_loss"])) < 1e-3 * float(fx["ctc_loss"])
    s, l, b = dino_ref.postprocess(logits, boxes, torch.tensor([[40.0, 1024.0]] * 2), 300)
    assert rel(s, fx["pp_scores"]) < 1e-6 and (l.numpy() == fx["pp_labels"]).all() and rel(b, fx["pp_boxes"]) < 1e-6


def test_training_mode_forward_quirk_Q3(golden_dir, sd):
    """train-mode forward with targets: 2*max_len DN queries stay in the output and feed the CTC loss."""
    fx = np.load(os.path.join(golden_dir, "dino_T_b2.npz"))
    cfg = dino_ref.default_cfg(num_queries=300)
    tg = synth.synth_targets(2, 166, seed=3)
    st = {}
    out = dino_ref.dino_forward(_sd(sd, 300), cfg, synth.synth_images(2, 40, 1024, seed=3), targets=tg, training=True, stages=st)
    assert out["pred_logits"].shape[1] == int(fx["pad_size"]) + 300
    assert rel(out["pred_logits"], fx["pred_logits"]) < TOL
    assert rel(out["pred_boxes"], fx["pred_boxes"]) < TOL
    loss, _ = dino_ref.loss_ctc(out["pred_logits"], out["pred_boxes"], tg)
    assert abs(loss.item() - float(fx["ctc_loss"])) < 1e-3 * float(fx["ctc_loss"])

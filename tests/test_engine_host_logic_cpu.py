"""CPU: the fused inference engine's host logic (dtlr_b200/engine.py) with its kernel launches replaced by the torch stand-ins of
tests/engine_ops_double.py, against the fixtures generated from the unmodified reference: BN folding and weight packing, level geometry,
buffer layouts and the launch order of the fp32 mode and of the split-precision mode (fp32 activations, [hi | hi | lo] x [hi | lo | hi]
products, split-output hand-over between layers, 16-bit attention core).  The CUDA kernels themselves are covered by the -m gpu tests."""
import numpy as np
import pytest
import torch

from dtlr_b200 import config, dino, synth
from dtlr_b200.misc import nested_tensor_from_tensor_list

import engine_ops_double as dbl
from gpu_common import fixture, rel


def _run(model, x, split, force=None, dtype=torch.float32):
    model.eval()
    model.compute_dtype = dtype
    model.split_precision = split
    model.use_cuda_graph = False
    model.transformer.debug_force_topk = force
    st = {}
    with torch.no_grad():
        out = model.engine().forward(nested_tensor_from_tensor_list(x), stages=st)
    model.transformer.debug_force_topk = None
    return out, st


def _build(num_queries):
    model, _, _ = dino.build_dino(config.latin_ctc_args(num_queries=num_queries))
    synth.load_synth_weights(model, 0)
    return model


@pytest.mark.parametrize("split", [False, True])
def test_engine_orchestration_config1_vs_reference_fixture(monkeypatch, split):
    dbl.install(monkeypatch)
    fx = fixture("dino_P_b1")
    model = _build(100)
    out, st = _run(model, synth.synth_images(1, 40, 704, seed=1), split, force=torch.from_numpy(fx["topk_idx"]).long() if split else None)
    tol = 1e-3
    if not split:
        assert (st["topk_idx"].numpy() == fx["topk_idx"]).all()
    for got, want in ((st["feats"][0][0][:, ::16, :, ::8], "feat_c3_s"), (st["feats"][2][0][:, ::32], "feat_c5_s"), (st["memory"][:, ::8, ::4], "memory_s"),
                      (st["topk_scores"], "topk_scores"), (st["hs"][0][:, ::8, ::4], "hs0_s"), (st["refs"][1], "ref1"), (out["pred_logits"], "pred_logits"),
                      (out["pred_boxes"], "pred_boxes"), (out["aux_outputs"][4]["pred_logits"][:, ::8, ::4], "aux4_logits_s"),
                      (out["aux_outputs"][0]["pred_boxes"], "aux0_boxes"), (out["interm_outputs"]["pred_logits"][:, ::8, ::4], "interm_logits_s"),
                      (out["interm_outputs"]["pred_boxes"], "interm_boxes"), (out["interm_outputs_for_matching_pre"]["pred_boxes"], "init_box_proposal")):
        e = rel(got.float(), fx[want])
        assert e < tol, (want, e)
    assert set(out) == {"pred_logits", "pred_boxes", "aux_outputs", "interm_outputs", "interm_outputs_for_matching_pre", "dn_meta"}
    c = dbl.CALLS
    if split:
        # every contraction is a split product: the 16 bottleneck 3x3 convs run on 3C-channel pixels, the exact-fp32 SIMT attention is not used
        # (width 704: the implicit-GEMM kernel cannot tile 176 / 88 / 44 / 22-pixel rows, so the 3x3 convs take the 16-bit im2col form)
        assert "conv2d_nhwc" not in c and c["im2col"] >= 16 and c["split_cast"] > 100 and c["cast"] == 6 and c["mha"] == 6
        assert out["pred_logits"].dtype == torch.float32 and st["memory"].dtype == torch.float32
    else:
        assert "split_cast" not in c and "conv2d_nhwc" not in c and c["mha"] == 6 and c["msda_fused"] == 12


def test_engine_orchestration_split_mode_config2_shape_implicit_convs(monkeypatch):
    """40 x 1024 / 900 queries (BASELINE config 2, batch 2): the widths the implicit-GEMM conv tiles (256 / 128 / 64 / 32-pixel rows), so the
    split-precision backbone runs conv1 -> (split output) -> implicit 3x3 over 3C-channel pixels -> (split output) -> conv3"""
    dbl.install(monkeypatch)
    fx = fixture("dino_A_b2")
    model = _build(900)
    out, st = _run(model, synth.synth_images(2, 40, 1024, seed=0), True, force=torch.from_numpy(fx["topk_idx"]).long())
    for got, want in ((st["feats"][2][0][:, ::32], "feat_c5_s"), (st["memory"][:, ::8, ::4], "memory_s"), (st["topk_scores"], "topk_scores"),
                      (out["pred_logits"], "pred_logits"), (out["pred_boxes"], "pred_boxes"), (out["interm_outputs"]["pred_boxes"], "interm_boxes")):
        e = rel(got.float(), fx[want])
        assert e < 1e-3, (want, e)
    # im2col: 16 inside the conv stand-in + the three strided 1x1 downsample convs (a gather of the fp32 map) + the stride-2 3x3 input_proj of the extra level
    assert dbl.CALLS["conv2d_nhwc"] == 16 and dbl.CALLS["im2col"] == 16 + 3 + 1


def test_engine_orchestration_ragged_batch_fp32(monkeypatch):
    """padded batch: level masks, valid ratios, masked value rows, proposals of padded tokens -- the non-cached geometry path"""
    dbl.install(monkeypatch)
    fx = fixture("dino_R_b3")
    model = _build(300)
    imgs = synth.synth_images(3, 40, 1024, seed=2, widths=fx["widths"].tolist())
    out, st = _run(model, imgs, False, force=torch.from_numpy(fx["topk_idx"]).long())
    for got, want in ((st["memory"][:, ::8, ::4], "memory_s"), (st["topk_scores"], "topk_scores"), (out["pred_logits"], "pred_logits"),
                      (out["pred_boxes"], "pred_boxes"), (out["interm_outputs"]["pred_boxes"], "interm_boxes")):
        e = rel(got.float(), fx[want])
        assert e < 1e-3, (want, e)
    assert dbl.CALLS["zero_masked_rows"] >= 7        # 6 encoder value projections + the decoder's batched one


def test_split_weight_pack_and_mode_switch(monkeypatch):
    """the engine keeps ONE packed weight set: switching fp32 -> split -> fp32 re-packs (distinct keys), 3x3 conv weights carry the
    per-tap [hi | lo | hi] layout, every other weight the whole-row one"""
    from dtlr_b200 import engine
    dbl.install(monkeypatch)
    model = _build(100)
    x = synth.synth_images(1, 40, 704, seed=1)
    _run(model, x, False)
    eng = model.engine()
    p32 = eng._packed
    assert p32["blocks"][0]["c2"][0].dtype == torch.float32 and p32["enc"][0]["l1"][0].shape == (2048, 256)
    _run(model, x, True)
    ps = eng._packed
    assert ps is not p32 and isinstance(eng._weight_dtype(), engine.SplitDtype)
    w2 = ps["blocks"][0]["c2"][0]
    assert w2.dtype == torch.float16 and w2.shape == (64, 9 * 3 * 64)
    t = w2.view(64, 9, 3, 64)
    assert torch.equal(t[:, :, 0], t[:, :, 2]) and not torch.equal(t[:, :, 0], t[:, :, 1])           # per tap: hi | lo | hi
    l1 = ps["enc"][0]["l1"][0]
    assert l1.shape == (2048, 768) and torch.equal(l1[:, :256], l1[:, 512:])                           # whole row: hi | lo | hi
    assert ps["tgt_embed"].dtype == torch.float32 and ps["stem_gemm"] is None
    model.split_precision = False
    assert eng._weight_dtype() == torch.float32


@pytest.mark.parametrize("dt,bound", [(torch.float16, 1e-2), (torch.bfloat16, 5e-2)])
def test_engine_orchestration_throughput_modes(monkeypatch, dt, bound):
    """the 16-bit modes bench.py's headline runs in: folded 16-bit conv weights, the tensor-core stem (staged im2col x [64, 152] GEMM), the fused
    FFN / box-head / Linear+LayerNorm launches, 16-bit token buffers -- host logic only, against the fixture within the modes' documented
    bounds (the stand-ins round operands and stored activations to 16 bits like the kernels)"""
    dbl.install(monkeypatch, half=True)
    fx = fixture("dino_A_b2")
    model = _build(900)
    out, st = _run(model, synth.synth_images(2, 40, 1024, seed=0), False, force=torch.from_numpy(fx["topk_idx"]).long(), dtype=dt)
    assert st["memory"].dtype == dt and out["pred_logits"].dtype == torch.float32 and out["pred_boxes"].dtype == torch.float32
    e_log, e_box = rel(out["pred_logits"], fx["pred_logits"]), rel(out["pred_boxes"], fx["pred_boxes"])
    assert e_log < bound and e_box < bound, (e_log, e_box)
    c = dbl.CALLS
    # 12 fused FFN blocks; box heads: the two-stage proposals + 6 decoder refinements + the batched shared head; 16 implicit 3x3 convs + 3 strided 1x1
    assert c["ffn_fused"] == 12 and c["mlp_head"] == 8 and c["conv2d_nhwc"] == 19 and "split_cast" not in c and "stem_conv" not in c

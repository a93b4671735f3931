"""CPU: the model boundary -- builder registry, state-dict layout, module sharing, config surface,
NestedTensor padding, decode helpers -- and the loud failure of the product path without CUDA."""
import json
import os

import pytest
import torch

from dtlr_b200 import config, dino, misc, synth
from dtlr_b200.registry import MODULE_BUILD_FUNCS


@pytest.fixture(scope="module")
def built():
    return MODULE_BUILD_FUNCS.get("dino")(config.latin_ctc_args())


def test_state_dict_layout_matches_reference(built, golden_dir):
    model = built[0]
    ref = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))
    sd = model.state_dict()
    assert set(sd) == set(ref)
    assert all(list(sd[k].shape) == ref[k] for k in ref)
    assert sum(p.numel() for p in model.parameters()) == 46951188            # SURVEY.md §6 [probed]
    assert sum(p.numel() for p in model.parameters() if p.requires_grad) == 46728788


def test_module_sharing_and_attribute_surface(built):
    model, criterion, post = built
    assert all(m is model.class_embed[0] for m in model.class_embed)         # dec_pred_class_embed_share
    assert all(m is model.bbox_embed[0] for m in model.bbox_embed)
    assert model.transformer.decoder.class_embed is model.class_embed
    assert model.transformer.enc_out_class_embed is not model.class_embed[0]  # two_stage_*_share=False -> deep copy
    assert model.class_embed[0].weight.data.shape[1] == 256                    # evaluation.py:60-86 reads this
    assert model.transformer.num_decoder_layers == 6 and model.dec_pred_class_embed_share
    assert model.label_enc.weight.shape == (168, 256)
    assert "loss_ce" in criterion.weight_dict and post["bbox"].num_select == 300
    assert any("backbone" in n for n, _ in model.named_parameters())         # util/get_param_dicts.py:25-28


def test_hwdb_config_head_width():
    model, _, _ = dino.build_dino(config.hwdb_args(num_queries=10))
    assert model.class_embed[0].weight.shape == (7356, 256) and model.label_enc.weight.shape == (7357, 256)
    assert model.dn_number == 0


def test_nested_tensor_padding():
    imgs = [torch.ones(3, 40, 64), torch.ones(3, 32, 100)]
    nt = misc.nested_tensor_from_tensor_list(imgs)
    assert nt.tensors.shape == (2, 3, 40, 100) and nt.mask.shape == (2, 40, 100)
    assert not nt.mask[0, :40, :64].any() and nt.mask[0, :, 64:].all() and nt.mask[1, 32:, :].all()
    assert nt.tensors[0, :, :, 64:].abs().sum() == 0


def test_ctc_view_branches_and_decode():
    logits = torch.full((1, 3, 4), -9.0)
    logits[0, 0, 2] = 9.0            # sum > 1-eps -> renormalised, class 2
    logits[0, 2, 1] = 0.2            # sum < 1: blank 1-s vs p -> p=0.55 > blank 0.45 -> class 1
    boxes = torch.tensor([[[0.9, 0, 0, 0], [0.5, 0, 0, 0], [0.1, 0, 0, 0]]])
    new = dino.ctc_view(logits, boxes)
    assert new.shape == (1, 3, 5)
    assert torch.allclose(new.sum(-1), torch.ones(1, 3), atol=1e-5)
    assert dino.convert_output_to_pred(new) == [[1, 2]]     # sorted by cx: query2 (class 1), query1 (blank), query0 (class 2)


def test_forward_refuses_cpu_tensors(built):
    model = built[0].eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.zeros(1, 3, 40, 64))


def test_synth_weights_are_key_determined():
    a = synth.synth_tensor("class_embed.3.weight", (166, 256))
    b = synth.synth_tensor("transformer.decoder.class_embed.0.weight", (166, 256))
    assert torch.equal(a, b)
    assert not torch.equal(a, synth.synth_tensor("transformer.enc_out_class_embed.weight", (166, 256)))


def test_collate_fn_like_reference():
    """util/misc.py:285-289: a list of (image, target) pairs -> (NestedTensor padded to the batch max, tuple of targets)"""
    import torch
    from dtlr_b200.misc import NestedTensor, collate_fn
    batch = [(torch.ones(3, 4, 7), {"labels": torch.tensor([1, 2])}), (torch.ones(3, 5, 6), {"labels": torch.tensor([3])})]
    samples, targets = collate_fn(batch)
    assert isinstance(samples, NestedTensor) and samples.tensors.shape == (2, 3, 5, 7) and isinstance(targets, tuple) and len(targets) == 2
    assert samples.mask[0, :4, :7].sum() == 0 and samples.mask[0, 4].all() and samples.mask[1, :, 6].all() and not samples.nopad
    assert targets[1]["labels"].tolist() == [3]


def test_pretrained_backbone_file_is_loaded_and_missing_file_is_loud(tmp_path):
    """reference backbone.py:118-121 starts from ImageNet weights (downloaded); here they come from a torchvision-format file"""
    import warnings
    from dtlr_b200 import backbone
    donor = backbone.ResNet50Body((1, 2, 3))
    sd = {k: torch.randn_like(v) for k, v in donor.state_dict().items()}
    tv = dict(sd)
    tv["fc.weight"], tv["fc.bias"] = torch.zeros(1000, 2048), torch.zeros(1000)        # dropped like IntermediateLayerGetter does
    tv["bn1.num_batches_tracked"] = torch.tensor(0)
    path = str(tmp_path / "resnet50.pth")
    torch.save(tv, path)
    bb = backbone.build_backbone(config.latin_ctc_args(backbone_pretrained=path))
    got = bb[0].body.state_dict()
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    with pytest.raises(FileNotFoundError):
        backbone.build_backbone(config.latin_ctc_args(backbone_pretrained=str(tmp_path / "nope.pth")))
    if backbone.find_pretrained_resnet50(config.latin_ctc_args()) is None:
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            backbone.build_backbone(config.latin_ctc_args())
        assert any("no ImageNet ResNet-50 weights" in str(x.message) for x in w)


def test_engine_cache_key_sees_replaced_modules_and_invalidate():
    """the packed-weight / graph cache key must change when a head is replaced (evaluation.py:60-86 style surgery)"""
    model = MODULE_BUILD_FUNCS.get("dino")(config.latin_ctc_args(num_queries=10))[0]
    eng = model.engine()
    k0 = eng._pack_key(torch.float32, "cpu")
    assert eng._pack_key(torch.float32, "cpu") == k0
    model.transformer.enc_out_class_embed = torch.nn.Linear(256, 166)          # fresh module: every _version is again 0/1
    k1 = eng._pack_key(torch.float32, "cpu")
    assert k1 != k0
    model.class_embed[0].weight.data[3] = 1.0                                  # invisible to autograd versions ...
    assert eng._pack_key(torch.float32, "cpu") == k1
    model.invalidate_engine()                                                  # ... so the caller says so
    assert eng._pack_key(torch.float32, "cpu") != k1
    k2 = eng._pack_key(torch.float32, "cpu")
    model.load_state_dict(model.state_dict())
    assert eng._pack_key(torch.float32, "cpu") != k2

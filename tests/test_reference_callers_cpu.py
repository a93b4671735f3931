"""CPU: the reference's OWN caller code, unmodified, driving dtlr_b200's model object ("attribute surgery by callers",
param-group selection).  Reads /root/reference, so it runs only in the development container (skipped where the tree is absent);
functions that live in scripts with heavy top-level side effects (evaluation.py) are extracted by AST and executed unmodified with
the globals they read."""
import ast
import copy
import os
import sys
import types

import pytest
import torch
import torch.nn as nn

REF = os.environ.get("DTLR_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models", "dino")), reason="reference tree not present")


def _build():
    from dtlr_b200 import config, dino, synth
    model, criterion, post = dino.build_dino(config.latin_ctc_args())
    synth.load_synth_weights(model, seed=0)
    return model


def _extract(path, name):
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name][0]
    return ast.Module(body=[fn], type_ignores=[])


def test_reference_load_model_class_head_surgery_runs_on_our_model(tmp_path):
    """evaluation.py:51-90 `load_model` with new_class_embedding: replaces model.class_embed (shared ModuleList),
    model.transformer.decoder.class_embed (a single Linear -- reference quirk), model.transformer.enc_out_class_embed and
    model.label_enc for a new charset, then strict-loads a checkpoint with those shapes."""
    new_charset = [chr(0x41 + i) for i in range(80)]
    # donor checkpoint: a model that went through the same surgery (what a fine-tuned checkpoint of the reference contains)
    donor = _build()
    d = donor.class_embed[0].weight.shape[1]
    torch.manual_seed(1)
    shared = nn.Linear(d, 80)
    donor.class_embed = nn.ModuleList([shared for _ in range(donor.transformer.num_decoder_layers)])
    donor.transformer.decoder.class_embed = nn.Linear(d, 80)
    donor.transformer.enc_out_class_embed = nn.Linear(d, 80)
    donor.label_enc = nn.Embedding(81, d)
    ckpt = os.path.join(tmp_path, "finetuned.pth")
    torch.save({"model": donor.state_dict()}, ckpt)

    model = _build()
    assert model.class_embed[0].out_features == 166
    ns = {"torch": torch, "nn": nn,
          "args": types.SimpleNamespace(new_class_embedding=True, fix_enc_out_class=False, new_label_enc=True, weights=ckpt),
          "args_dataset": types.SimpleNamespace(device="cpu", charset=new_charset),
          "dataset_val": types.SimpleNamespace(charset=new_charset)}
    exec(compile(_extract(os.path.join(REF, "evaluation.py"), "load_model"), "evaluation.py", "exec"), ns)
    ns["load_model"](model)                       # mutates in place (the surgery branch of the reference returns None)

    assert not model.training
    assert all(ce is model.class_embed[0] for ce in model.class_embed) and model.class_embed[0].out_features == 80
    assert isinstance(model.transformer.decoder.class_embed, nn.Linear) and model.transformer.decoder.class_embed.out_features == 80
    assert model.transformer.enc_out_class_embed.out_features == 80 and model.label_enc.num_embeddings == 81
    sd, want = model.state_dict(), donor.state_dict()
    assert sd.keys() == want.keys() and all(torch.equal(sd[k], want[k]) for k in sd)
    # the plain branch (no surgery): strict load of a stock checkpoint
    stock = _build()
    torch.save({"model": stock.state_dict()}, ckpt)
    ns["args"].new_class_embedding = False
    other = _build()
    with torch.no_grad():
        other.class_embed[0].bias.add_(1.0)
    got = ns["load_model"](other)
    assert got is other and torch.equal(got.class_embed[0].bias, stock.class_embed[0].bias)


def test_reference_param_groups_on_our_model():
    """util/get_param_dicts.py:15-34 (`"backbone" in name` split, finetuning.py optimizer set-up): same parameter counts as the
    reference model's layout -- 2 groups, backbone group carries lr_backbone."""
    sys.path.insert(0, REF)
    try:
        from util.get_param_dicts import get_param_dict
    finally:
        sys.path.remove(REF)
    model = _build()
    groups = get_param_dict(types.SimpleNamespace(lr_backbone=1e-5), model)
    assert len(groups) == 2 and groups[1]["lr"] == 1e-5
    n_other = sum(p.numel() for p in groups[0]["params"])
    n_backbone = sum(p.numel() for p in groups[1]["params"])
    assert n_other + n_backbone == 46728788                     # trainable parameters of the reference model
    assert n_backbone == sum(p.numel() for n, p in model.named_parameters() if n.startswith("backbone.") and p.requires_grad) > 0
    opt = torch.optim.AdamW(groups, lr=1e-4, weight_decay=1e-4)  # finetuning.py builds exactly this
    assert len(opt.param_groups) == 2

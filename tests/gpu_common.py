"""helpers shared by the -m gpu parity tests"""
import json
import os

import numpy as np
import torch

from dtlr_b200 import config, dino, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    fin = torch.isfinite(a) & torch.isfinite(b)
    assert (torch.isfinite(a) == torch.isfinite(b)).all(), "non-finite pattern differs"
    return ((a - b)[fin].abs().max() / b[fin].abs().max().clamp_min(1e-12)).item()


def fixture(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def build_model(num_queries, seed=0, **over):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model, crit, post = dino.build_dino(config.latin_ctc_args(num_queries=num_queries, **over))
    synth.load_synth_weights(model, seed)
    return model.cuda(), crit, post


def near_tie_mask(fx):
    """positions of the reference ranking whose score is within 1e-4 of a neighbour (ranking numerically undecidable)"""
    ref_scores = np.take_along_axis(fx["topk_scores"], fx["topk_idx"].astype(np.int64), 1)
    return np.minimum(np.abs(np.diff(ref_scores, axis=1, prepend=np.inf)),
                      np.abs(np.diff(ref_scores, axis=1, append=-np.inf))) < 1e-4

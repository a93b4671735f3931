"""Deterministic synthetic weights and inputs for parity tests, golden fixtures and bench.py.

There are no checkpoints or images in the build/run environment (SURVEY.md §7 hard part 3), and a
47 M-parameter state_dict cannot be committed as a fixture.  So every tensor of a DINO state_dict is
regenerated from *its key name and shape alone*: generator seed = crc32(canonical key) ^ seed, drawn on
the CPU generator (bit-reproducible across hosts for one torch version).  The same function is used
 * by tests/golden/make_golden.py to fill the UNMODIFIED reference model before dumping outputs,
 * by the parity tests to fill the oracle and the CUDA model on the GPU box,
 * by bench.py (random-init weights of the named architecture).

The distribution is *not* the reference init on purpose (SURVEY.md §7 hard part 2, "test-weight caveat"):
the reference zero-inits the last bbox-MLP layer and uses class bias -4.6, which makes boxes insensitive
to everything and the CTC decode degenerate.  Here every layer is random, the class bias is lowered so
blank and non-blank frames both occur, and the two-stage class head is widened so top-k scores are
well separated.
"""
import math
import re
import zlib

import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)

_ALIAS = [
    # dec_pred_{class,bbox}_embed_share=True: six ModuleList slots + the decoder aliases are ONE module
    # (reference models/dino/dino.py:170-191) -> canonical name "<kind>_embed.0".
    (re.compile(r"^(?:transformer\.decoder\.)?(class_embed|bbox_embed)\.\d+\."), r"\1.0."),
]


def canonical_key(key: str) -> str:
    for pat, rep in _ALIAS:
        key = pat.sub(rep, key)
    return key


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(canonical_key(key).encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synth_tensor(key: str, shape, seed: int = 0, n_heads: int = 8, n_levels: int = 4,
                 n_points: int = 4) -> torch.Tensor:
    """fp32 CPU tensor for one state_dict entry, determined by (key, shape, seed)."""
    shape = tuple(int(s) for s in shape)
    g = _gen(key, seed)
    ck = canonical_key(key)
    leaf = ck.rsplit(".", 1)[-1]

    def randn(std=1.0):
        return torch.randn(shape, generator=g, dtype=torch.float32) * std

    def rand():
        return torch.rand(shape, generator=g, dtype=torch.float32)

    # ---- frozen batch-norm buffers of the ResNet body
    if ".bn" in ck or "downsample.1." in ck:
        if leaf == "weight":
            return 1.0 + 0.1 * randn()
        if leaf == "bias":
            return 0.05 * randn()
        if leaf == "running_mean":
            return 0.05 * randn()
        if leaf == "running_var":
            return 0.8 + 0.4 * rand()
    # ---- LayerNorm / GroupNorm affine
    if re.search(r"(norm\d?|enc_output_norm|input_proj\.\d\.1)\.(weight|bias)$", ck):
        return 1.0 + 0.1 * randn() if leaf == "weight" else 0.05 * randn()
    # ---- MSDeformAttn: keep the reference's directional grid bias, add noise so offsets are data dependent
    if ck.endswith("sampling_offsets.bias"):
        thetas = torch.arange(n_heads, dtype=torch.float32) * (2.0 * math.pi / n_heads)
        grid = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(n_heads, 1, 1, 2).repeat(1, n_levels, n_points, 1)
        for i in range(n_points):
            grid[:, :, i, :] *= i + 1
        return grid.reshape(-1)[: shape[0]] + 0.3 * randn()
    if ck.endswith("sampling_offsets.weight"):
        return randn(0.04)
    if ck.endswith("attention_weights.weight"):
        return randn(0.06)
    if ck.endswith("attention_weights.bias"):
        return randn(0.3)
    # ---- heads
    if ck.endswith("enc_out_class_embed.weight"):
        return randn(0.5)          # wide two-stage scores -> well separated top-k ranking
    if ck.endswith("class_embed.0.weight"):
        return randn(0.12)
    if ck.endswith("class_embed.bias") or ck.endswith("class_embed.0.bias"):
        return -7.3 + 0.5 * randn()   # sum_c sigmoid straddles 1 -> blank and non-blank frames
    if re.search(r"bbox_embed(\.0)?\.layers\.2\.weight$", ck):
        return randn(0.03)
    if re.search(r"bbox_embed(\.0)?\.layers\.2\.bias$", ck):
        return randn(0.05)
    # ---- embeddings / level embed
    if ck.endswith("tgt_embed.weight") or ck.endswith("level_embed") or ck.endswith("label_enc.weight"):
        return randn(1.0 if "tgt_embed" in ck else 0.5)
    # ---- generic conv / linear
    if leaf in ("weight", "in_proj_weight") and len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        # He-style for the ReLU conv trunk (keeps activations O(1) through 50 layers), Xavier-ish elsewhere
        if len(shape) == 4:
            std = math.sqrt(2.0 / fan_in) * (0.45 if re.search(r"conv3\.weight$|downsample\.0\.weight$", ck) else 1.0)
        else:
            std = math.sqrt(1.0 / fan_in)
        return randn(std)
    if leaf in ("bias", "in_proj_bias"):
        return randn(0.02)
    return randn(0.02)


def synth_state_dict(shapes: dict, seed: int = 0) -> dict:
    """shapes: {key: shape}.  Returns {key: fp32 CPU tensor}; aliased keys get identical values."""
    return {k: synth_tensor(k, s, seed) for k, s in shapes.items()}


def load_synth_weights(model: torch.nn.Module, seed: int = 0) -> None:
    """Fill any nn.Module whose state_dict follows the reference key layout."""
    sd = model.state_dict()
    new = synth_state_dict({k: v.shape for k, v in sd.items()}, seed)
    model.load_state_dict({k: new[k].to(v.dtype) for k, v in sd.items()})


def synth_images(batch: int, height: int = 40, width: int = 1024, seed: int = 0, widths=None):
    """Grayscale text-line stand-ins: g = rand(B,1,H,W) replicated to RGB
    (reference datasets/IAM.py:86-88 convert("RGB")) and ImageNet-normalised (datasets/IAM.py:110-112).
    widths: optional per-image widths (ragged batch) -> returns a list of (3,H,w_i) tensors.
    Returns fp32 CPU tensor (B,3,H,W) or a list."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * seed + 17)
    mean = torch.tensor(IMAGENET_MEAN).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(3, 1, 1)
    if widths is None:
        gray = torch.rand((batch, 1, height, width), generator=g)
        return (gray.expand(-1, 3, -1, -1) - mean) / std
    out = []
    for w in widths:
        gray = torch.rand((1, height, int(w)), generator=g)
        out.append((gray.expand(3, -1, -1) - mean) / std)
    return out


def synth_targets(batch: int, num_classes: int, seed: int = 0, mean_len: float = 42.0, std_len: float = 9.0,
                  min_len: int = 5, max_len: int = 91):
    """CTC targets with IAM-like length statistics: labels randint(0,C), boxes zeros
    (reference datasets/IAM.py:100-102)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(7919 * seed + 3)
    targets = []
    for _ in range(batch):
        n = int(torch.clamp(torch.round(torch.randn((), generator=g) * std_len + mean_len), min_len, max_len))
        labels = torch.randint(0, num_classes, (n,), generator=g)
        targets.append({"labels": labels, "boxes": torch.zeros(n, 4)})
    return targets

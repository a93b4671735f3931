"""Precision simulator (test infrastructure): the CPU oracle (oracle/dino_ref.py) re-run with the operand / storage roundings of
a candidate GPU compute mode, so that the error budget of a mode (which tensors may be bf16 / fp16 / split, which must stay
fp32) is known BEFORE kernels are written, and so that the per-stage budget table in DESIGN.md can be regenerated.

    python -m tests.precision_sim [--batch 2] [--width 1024] [--queries 900]

A mode is (op_bits, store_bits, score_bits): significand bits (implicit one included) of the MMA operands, of the activations
written to HBM between kernels, and of the operands on the two-stage score path (backbone -> encoder -> enc_output -> class max).
8 = bf16, 11 = fp16, 16 = 2xbf16 split (hi + lo, 3 MMAs), 22 = 2xfp16 split, 24 = fp32.  Exponent range is not modelled.
"""
import argparse
import math
import re
import sys
import types

import torch
import torch.nn.functional as TF

sys.path.insert(0, __file__.rsplit("/tests/", 1)[0])
from dtlr_b200 import config, dino, synth  # noqa: E402
from oracle import dino_ref  # noqa: E402


def round_bits(x, bits):
    if bits >= 24 or not torch.is_floating_point(x):
        return x
    if bits == 8:
        return x.to(torch.bfloat16).to(x.dtype)
    if bits == 11:
        return x.to(torch.float16).to(x.dtype)
    m, e = torch.frexp(x)
    s = float(1 << bits)
    return torch.ldexp(torch.round(m * s) / s, e)


class Mode:
    """rules: list of (regex on the site name, op_bits, store_bits); first match wins; default fp32.  A site name is the
    state-dict prefix of the weight an op uses ("transformer.decoder.layers.3.linear1"), or "<layer prefix>.sa.qk|sa.p|sa.v"
    for the attention contractions, "msda.core:<early|late>" for the deformable-attention core and its offsets / weights."""

    def __init__(self, name, rules=(), default=(24, 24)):
        self.name, self.rules, self.default = name, [(re.compile(r), o, s) for r, o, s in rules], default
        self.late = False
        self.cache = {}

    def bits(self, site):
        if site not in self.cache:
            self.cache[site] = next(((o, s) for r, o, s in self.rules if r.search(site)), self.default)
        return self.cache[site]

    def o(self, x, site):
        return round_bits(x, self.bits(site)[0])

    def s(self, x, site):
        return round_bits(x, self.bits(site)[1])


class _FProxy(types.ModuleType):
    """torch.nn.functional with linear / conv2d / layer_norm / group_norm rounding their operands and results as `mode` says."""

    def __init__(self, mode, names):
        super().__init__("F_proxy")
        self.mode, self.names = mode, names

    def __getattr__(self, k):
        return getattr(TF, k)

    def site(self, w):
        return self.names.get(id(w), "?")

    def linear(self, x, w, b=None, site=None):
        md, site = self.mode, site or self.site(w)
        return md.s(TF.linear(md.o(x, site), md.o(w, site), b), site)

    def conv2d(self, x, w, b=None, **kw):
        md, site = self.mode, self.site(w)
        y = TF.conv2d(md.o(x, site), md.o(w, site), b, **kw)
        return y if "backbone" in site else md.s(y, site)     # FrozenBN follows in fp32 (epilogue); stored after the ReLU

    def layer_norm(self, x, shp, w, b, eps):
        return self.mode.s(TF.layer_norm(x, shp, w, b, eps), self.site(w))

    def group_norm(self, x, g, w, b, eps):
        return self.mode.s(TF.group_norm(x, g, w, b, eps), self.site(w))

    def relu(self, x):
        return self.mode.s(TF.relu(x), "relu:" + ("late" if self.mode.late else "early"))


def run(mode, sd, cfg, x, force_topk=None):
    """oracle forward under `mode`; returns (out, stages)."""
    names = {}
    for k, v in sd.items():
        names[id(v)] = k.rsplit(".", 1)[0]
    Fp = _FProxy(mode, names)
    saved = (dino_ref.F, dino_ref.mha_self_attention, dino_ref._msda.msda_forward)
    mode.late = False

    def mha(sd_, p, qk_in, v_in, nheads, attn_mask=None):
        B, Q, C = qk_in.shape
        Wi, bi = sd_[p + ".in_proj_weight"], sd_[p + ".in_proj_bias"]
        q = Fp.linear(qk_in, Wi[:C], bi[:C], site=p + ".in_proj.q")
        k = Fp.linear(qk_in, Wi[C:2 * C], bi[C:2 * C], site=p + ".in_proj.k")
        v = Fp.linear(v_in, Wi[2 * C:], bi[2 * C:], site=p + ".in_proj.v")
        dh = C // nheads
        q, k, v = (t.view(B, Q, nheads, dh).transpose(1, 2) for t in (q, k, v))
        s = (mode.o(q, p + ".sa.qk") * (1.0 / math.sqrt(dh))) @ mode.o(k, p + ".sa.qk").transpose(-1, -2)
        pr = mode.o(TF.softmax(s, -1), p + ".sa.p")
        o = mode.s((pr @ mode.o(v, p + ".sa.v")).transpose(1, 2).reshape(B, Q, C), p + ".sa.v")
        return dino_ref._linear(sd_, p + ".out_proj", o)

    core = saved[2]

    def msda_forward(value, shapes, lsi, loc, aw):
        site = "msda.core:" + ("late" if mode.late else "early")
        return mode.s(core(mode.o(value, site), shapes, lsi, loc, aw), site)

    orig_topk, orig_gather = torch.topk, torch.gather

    def topk_hook(*a, **k):
        mode.late = True       # everything after the two-stage top-k is "late"
        return orig_topk(*a, **k)

    def gather_hook(*a, **k):  # the oracle skips torch.topk when the ranking is forced; the gathers follow it either way
        mode.late = True
        return orig_gather(*a, **k)

    dino_ref.F = Fp
    dino_ref.mha_self_attention = mha
    dino_ref._msda.msda_forward = msda_forward
    torch.topk, torch.gather = topk_hook, gather_hook
    try:
        st = {}
        out = dino_ref.dino_forward(sd, cfg, x, stages=st, force_topk=force_topk)
    finally:
        dino_ref.F, dino_ref.mha_self_attention, dino_ref._msda.msda_forward = saved
        torch.topk, torch.gather = orig_topk, orig_gather
    return out, st


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


def frames_of(out, eps=0.003):
    new, _ = dino_ref.ctc_view(out["pred_logits"], out["pred_boxes"], eps)
    return new.argmax(-1)


GROUPS = [          # (name, regex) -- a partition of the op sites of the forward, in path order
    ("backbone convs", r"^backbone|^relu:early"),
    ("input_proj", r"^input_proj"),
    ("enc attn projections", r"encoder\.layers\.\d\.self_attn"),
    ("enc msda core", r"^msda\.core:early"),
    ("enc ffn + norms", r"encoder\.layers\.\d\.(linear|norm)"),
    ("score head (enc_output, class, bbox)", r"transformer\.enc_out"),
    ("dec ref_point_head", r"ref_point_head"),
    ("dec sa projections", r"decoder\.layers\.\d\.self_attn\.(in_proj|out_proj)"),
    ("dec sa core (QK, P, V)", r"decoder\.layers\.\d\.self_attn\.sa\."),
    ("dec ca projections", r"decoder\.layers\.\d\.cross_attn"),
    ("dec msda core", r"^msda\.core:late"),
    ("dec ffn + norms", r"decoder\.layers\.\d\.(linear|norm)|decoder\.norm|^relu:late"),
    ("dec bbox refine MLP", r"decoder\.bbox_embed"),
    ("heads (bbox, class)", r"^bbox_embed|^class_embed"),
]
EARLY = GROUPS[:6]
LATE = GROUPS[6:]


def _all(groups, o, s):
    return [(rx, o, s) for _, rx in groups]


MODES = {
    "bf16 op + bf16 store (round-1 engine)": _all(GROUPS, 8, 8),
    "bf16 op, fp32 store": _all(GROUPS, 8, 24),
    "fp16 op + fp16 store": _all(GROUPS, 11, 11),
    "fp16 op, fp32 store": _all(GROUPS, 11, 24),
    "2xbf16 split op, fp32 store": _all(GROUPS, 16, 24),
    "2xbf16 early / fp16 late, fp32 store": _all(EARLY, 16, 24) + _all(LATE, 11, 24),
    "2xfp16 split op, fp32 store": _all(GROUPS, 22, 24),
    # the engine's split-precision mode (model.split_precision): every Linear / conv as a 2xfp16 split product, the deformable-attention
    # core in exact fp32, the decoder self-attention core on fp16 operands with an fp16 result, activations fp32
    "split engine: 2xfp16 GEMM, fp16 sa core": [(r"self_attn\.sa\.", 11, 11), (r"^msda\.core", 24, 24)] + _all(GROUPS, 22, 24),
    "split engine, bf16 halves": [(r"self_attn\.sa\.", 8, 8), (r"^msda\.core", 24, 24)] + _all(GROUPS, 16, 24),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--queries", type=int, default=900)
    ap.add_argument("--modes", default="")
    ap.add_argument("--sens", type=int, default=0, help="per-group sensitivity: this many operand bits in ONE group, fp32 elsewhere")
    ap.add_argument("--sens-store", type=int, default=24)
    a = ap.parse_args()
    torch.manual_seed(0)
    model, _, _ = dino.build_dino(config.latin_ctc_args(num_queries=a.queries))
    synth.load_synth_weights(model, 0)
    sd = {k: v.detach().float() for k, v in model.state_dict().items()}
    cfg = dino_ref.default_cfg(num_queries=a.queries)
    x = synth.synth_images(a.batch, 40, a.width, seed=0)
    ref, rst = run(Mode("fp32"), sd, cfg, x)
    rframes = frames_of(ref)
    rs = rst["topk_scores"]
    gaps = torch.gather(rs, 1, rst["topk_idx"])
    gaps = (gaps[:, :-1] - gaps[:, 1:])
    print("reference: score range [%.3f, %.3f], median rank gap %.2e, gaps<1e-5: %.2f%%, <1e-4: %.2f%%" % (
        rs.min(), rs.max(), gaps.median(), 100 * (gaps < 1e-5).float().mean(), 100 * (gaps < 1e-4).float().mean()))
    print("%-40s %9s %9s %9s %9s | %9s %9s %9s | %8s" % (
        "mode (ranking forced)", "feat_c5", "memory", "scores", "hs5", "logits", "boxes", "frames!=", "rank=="))
    modes = dict(MODES)
    if a.sens:
        modes = {"%s @%d/%d bits" % (n, a.sens, a.sens_store): [(rx, a.sens, a.sens_store)] for n, rx in GROUPS}
    for name, rules in modes.items():
        if a.modes and not any(m in name for m in a.modes.split(",")):
            continue
        md = Mode(name, rules)
        out, st = run(md, sd, cfg, x, force_topk=rst["topk_idx"])
        fr = frames_of(out)
        rank = torch.topk(st["topk_scores"], a.queries, dim=1)[1]
        print("%-40s %9.2e %9.2e %9.2e %9.2e | %9.2e %9.2e %9.5f | %8.4f" % (
            name[:40], rel(st["feats"][-1], rst["feats"][-1]), rel(st["memory"], rst["memory"]),
            (st["topk_scores"] - rs).abs().max().item(), rel(st["hs"][-1], rst["hs"][-1]),
            rel(out["pred_logits"], ref["pred_logits"]), rel(out["pred_boxes"], ref["pred_boxes"]),
            (fr != rframes).float().mean().item(), (rank == rst["topk_idx"]).float().mean().item()))
        sys.stdout.flush()


if __name__ == "__main__":
    main()

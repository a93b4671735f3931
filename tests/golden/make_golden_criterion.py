"""Golden vectors of the detection loss: the UNMODIFIED reference SetCriterion.forward_standard +
HungarianMatcher (scipy) run on CPU on small random detection outputs.  Run in the build container:

    python tests/golden/make_golden_criterion.py

Inputs are stored in the fixture (they are tiny) together with the reference losses and matcher indices."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shims  # noqa: E402

warnings.filterwarnings("ignore")


def make_outputs(g, B, Q, C, n_aux, sizes, dn_groups):
    def det(q):
        return {"pred_logits": torch.randn(B, q, C, generator=g) * 2 - 2,
                "pred_boxes": torch.cat([torch.rand(B, q, 2, generator=g), torch.rand(B, q, 2, generator=g) * 0.2 + 0.01], -1)}
    out = det(Q)
    out["aux_outputs"] = [det(Q) for _ in range(n_aux)]
    out["interm_outputs"] = det(Q)
    targets = []
    for n in sizes:
        targets.append({"labels": torch.randint(0, C, (n,), generator=g),
                        "boxes": torch.cat([torch.rand(n, 2, generator=g), torch.rand(n, 2, generator=g) * 0.2 + 0.01], -1)})
    if dn_groups:
        single = max(sizes) * 2          # positive + negative half per group (reference prepare_for_cdn)
        pad = single * dn_groups
        known = det(pad)
        known["aux_outputs"] = [det(pad) for _ in range(n_aux)]
        out["dn_meta"] = {"pad_size": pad, "num_dn_group": dn_groups, "output_known_lbs_bboxes": known}
    else:
        out["dn_meta"] = None
    return out, targets


def flatten(prefix, d, dst):
    dst[prefix + "pred_logits"] = d["pred_logits"].numpy()
    dst[prefix + "pred_boxes"] = d["pred_boxes"].numpy()


def main():
    ref_shims.load_reference()
    _to = torch.Tensor.to

    def to(self, *a, **k):                                        # the reference hard-codes .to("cuda") for the zero DN losses
        if a and isinstance(a[0], str) and a[0].startswith("cuda"):
            return self
        return _to(self, *a, **k)
    torch.Tensor.to = to
    from models.dino.dino import SetCriterion
    from models.dino.matcher import HungarianMatcher
    cases = {"D1": dict(B=3, Q=40, C=12, n_aux=2, sizes=[5, 0, 9], dn_groups=0, train=False),
             "D2": dict(B=2, Q=64, C=20, n_aux=1, sizes=[7, 3], dn_groups=2, train=True),
             "D3": dict(B=2, Q=300, C=97, n_aux=0, sizes=[60, 100], dn_groups=0, train=True)}
    fx = {}
    for name, c in cases.items():
        g = torch.Generator().manual_seed(hash(name) % 1000 if False else {"D1": 11, "D2": 22, "D3": 33}[name])
        out, targets = make_outputs(g, c["B"], c["Q"], c["C"], c["n_aux"], c["sizes"], c["dn_groups"])
        matcher = HungarianMatcher(cost_class=2.0, cost_bbox=5.0, cost_giou=2.0, focal_alpha=0.25)
        crit = SetCriterion(c["C"], matcher=matcher, weight_dict={}, focal_alpha=0.25, losses=["labels", "boxes", "cardinality"])
        crit.train(c["train"])
        losses, ind = crit.forward_standard(out, targets, return_indices=True)
        p = name + "/"
        fx[p + "meta"] = np.array([c["B"], c["Q"], c["C"], c["n_aux"], c["dn_groups"], int(c["train"])], dtype=np.int64)
        fx[p + "sizes"] = np.array(c["sizes"], dtype=np.int64)
        flatten(p + "out/", out, fx)
        for i, a in enumerate(out["aux_outputs"]):
            flatten(p + "aux%d/" % i, a, fx)
        flatten(p + "interm/", out["interm_outputs"], fx)
        if c["dn_groups"]:
            k = out["dn_meta"]["output_known_lbs_bboxes"]
            flatten(p + "dn/", k, fx)
            for i, a in enumerate(k["aux_outputs"]):
                flatten(p + "dn_aux%d/" % i, a, fx)
            fx[p + "pad_size"] = np.array(out["dn_meta"]["pad_size"])
        for b, t in enumerate(targets):
            fx[p + "tgt%d/labels" % b] = t["labels"].numpy()
            fx[p + "tgt%d/boxes" % b] = t["boxes"].numpy()
        for k, v in losses.items():
            fx[p + "loss/" + k] = np.asarray(float(v), dtype=np.float64)
        # indices_list order: aux 0..n-1, interm, final  (reference dino.py:871-962)
        for li, layer in enumerate(ind):
            for b, (i, j) in enumerate(layer):
                fx[p + "ind%d/%d/src" % (li, b)] = np.asarray(i, dtype=np.int64)
                fx[p + "ind%d/%d/tgt" % (li, b)] = np.asarray(j, dtype=np.int64)
        print(name, {k: round(float(v), 5) for k, v in list(losses.items())[:8]}, "...", len(losses), "losses")
    np.savez_compressed(os.path.join(HERE, "criterion.npz"), **fx)
    print("wrote criterion.npz", os.path.getsize(os.path.join(HERE, "criterion.npz")), "bytes")


if __name__ == "__main__":
    main()

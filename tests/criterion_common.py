"""Loads tests/golden/criterion.npz (reference SetCriterion.forward_standard outputs, see make_golden_criterion.py)."""
import os

import numpy as np
import torch

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "criterion.npz")
CASES = ("D1", "D2", "D3")


def load_case(name, device="cpu"):
    z = np.load(PATH)
    p = name + "/"
    B, Q, C, n_aux, dn_groups, train = [int(v) for v in z[p + "meta"]]
    sizes = [int(v) for v in z[p + "sizes"]]

    def det(prefix):
        return {"pred_logits": torch.from_numpy(z[p + prefix + "pred_logits"]).to(device),
                "pred_boxes": torch.from_numpy(z[p + prefix + "pred_boxes"]).to(device)}
    out = det("out/")
    out["aux_outputs"] = [det("aux%d/" % i) for i in range(n_aux)]
    out["interm_outputs"] = det("interm/")
    if dn_groups:
        known = det("dn/")
        known["aux_outputs"] = [det("dn_aux%d/" % i) for i in range(n_aux)]
        out["dn_meta"] = {"pad_size": int(z[p + "pad_size"]), "num_dn_group": dn_groups, "output_known_lbs_bboxes": known}
    else:
        out["dn_meta"] = None
    targets = [{"labels": torch.from_numpy(z[p + "tgt%d/labels" % b]).to(device),
                "boxes": torch.from_numpy(z[p + "tgt%d/boxes" % b]).to(device)} for b in range(B)]
    losses = {k[len(p + "loss/"):]: float(z[k]) for k in z.files if k.startswith(p + "loss/")}
    # reference indices_list order: aux 0..n-1, interm, final
    indices = []
    for li in range(n_aux + 2):
        indices.append([(torch.from_numpy(z[p + "ind%d/%d/src" % (li, b)]), torch.from_numpy(z[p + "ind%d/%d/tgt" % (li, b)]))
                        for b in range(B)])
    return dict(B=B, Q=Q, C=C, n_aux=n_aux, dn_groups=dn_groups, train=bool(train), sizes=sizes, outputs=out, targets=targets,
                losses=losses, indices=indices)

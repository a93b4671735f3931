"""Golden vectors for the stages either side of the forward, produced by the UNMODIFIED reference:

    python tests/golden/make_golden_io.py        (build container only: needs /root/reference, PIL, torchvision)

  * datasets/transforms.py  RandomResize -> ToTensor -> Normalize, then util/misc.py nested_tensor_from_tensor_list, on
    synthetic grayscale "line" images (PIL mode L converted to RGB as datasets/IAM.py:86-88 does).  The fixture stores the
    RESIZED u8 pixels (what our GPU stage takes), the reference's fp32 batch + mask, and the size rule on a table of sizes;
  * ngram/prediction_helpers.py get_new_pred_logits (torchaudio stubbed -- imported, never called; the hard-coded
    `.to("cuda")` mapped to the CPU) for multipliers 1 and 2, both blank branches exercised;
  * evaluation.py metric functions, extracted from the source by AST (the script body needs a dataset and a GPU) and executed
    unmodified on string / word-list pairs.
Writes tests/golden/io.npz and tests/golden/io_metrics.json.
"""
import ast
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shims  # noqa: E402

REF = ref_shims.REF_ROOT


def transforms_case():
    from PIL import Image
    ref_shims.load_reference()
    import datasets.transforms as T          # reference module
    from util.misc import nested_tensor_from_tensor_list
    rng = np.random.default_rng(0)
    norm = T.Compose([T.ToTensor(), T.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])])
    tf = T.Compose([T.RandomResize([40], max_size=512), norm])          # eval transform shape of datasets/IAM.py:225-229
    out = {}
    tensors = []
    for i, (h, w) in enumerate([(57, 513), (40, 256), (61, 1211), (33, 151)]):
        img = Image.fromarray(rng.integers(0, 256, (h, w), dtype=np.uint8), mode="L").convert("RGB")
        resized, _ = T.RandomResize([40], max_size=512)(img, None)
        out["img%d_u8" % i] = np.asarray(resized)[:, :, 0].copy()          # R = G = B
        t, _ = norm(resized, None)
        tensors.append(t)
        t2, _ = tf(img, None)
        assert torch.equal(t, t2)
    # one true-RGB image batch as well
    rgb = [Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), mode="RGB") for h, w in [(40, 100), (36, 77)]]
    rgb_t = [norm(im, None)[0] for im in rgb]
    for i, im in enumerate(rgb):
        out["rgb%d_u8" % i] = np.asarray(im).copy()
    nt = nested_tensor_from_tensor_list(tensors)
    out["batch"], out["mask"] = nt.tensors.numpy(), nt.mask.numpy()
    nt = nested_tensor_from_tensor_list(rgb_t)
    out["rgb_batch"], out["rgb_mask"] = nt.tensors.numpy(), nt.mask.numpy()
    # the size rule, called through the reference's resize() on blank images
    table = []
    for (w, h, size, mx) in [(913, 57, 40, 1024), (512, 40, 40, 1024), (2011, 61, 40, 1024), (301, 33, 40, 1024), (1333, 94, 800, 1333),
                             (2200, 128, 800, 1333), (100, 300, 800, 1333), (800, 800, 800, 1333), (640, 480, 800, 1333)]:
        r, _ = T.resize(Image.new("L", (w, h)), None, size, mx)
        table.append([w, h, size, mx, r.size[1], r.size[0]])
    out["size_table"] = np.array(table, dtype=np.int64)
    return out


def ngram_case():
    sys.modules.setdefault("torchaudio", types.ModuleType("torchaudio"))
    for name in ("torchaudio.models", "torchaudio.models.decoder"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["torchaudio.models.decoder"].ctc_decoder = None
    sys.path.insert(0, os.path.join(REF, "ngram"))
    import prediction_helpers as ph            # reference module
    real_to = torch.Tensor.to
    torch.Tensor.to = lambda self, *a, **k: self if (a and a[0] == "cuda") else real_to(self, *a, **k)
    try:
        g = torch.Generator().manual_seed(3)
        logits = torch.randn(2, 50, 20, generator=g) * 1.5 - 6.0
        logits[:, ::3] += 5.0                                   # rows whose probabilities sum above 1 (second branch)
        boxes = torch.rand(2, 50, 4, generator=g)
        o = {"np_logits": logits.numpy(), "np_boxes": boxes.numpy()}
        for mult in (1, 2):
            o["new_pred_x%d" % mult] = ph.get_new_pred_logits({"pred_logits": logits, "pred_boxes": boxes}, mult).numpy()
    finally:
        torch.Tensor.to = real_to
    return o


def metric_case():
    src = open(os.path.join(REF, "evaluation.py")).read()
    tree = ast.parse(src)
    want = {"character_error_rate", "word_error_rate", "split_labels_into_words", "process_pred_string"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    for n in body:
        n.decorator_list = []                  # @torch.no_grad() on a pure-python function
    ns = {"re": __import__("re")}
    exec(compile(ast.Module(body=body, type_ignores=[]), "evaluation.py", "exec"), ns)
    charset = list(" abcdefghijklmnopqrstuvwxyzBCITV.,-'0123456789€")
    pairs = [("the quick brown fox", "the quick brown fox"), ("the quick brown fox", "teh quik brown  fx"), ("", "abc"), ("abc", ""),
             ("B B C news , at 10 - 12 . 'quoted ' 1, 2", "BBC news, at 10-12. 'quoted' 1,2"), ("a..b ...c ,,d 5€x", "a.b ...c ,d 5 € x"),
             ("kitten", "sitting"), ("flaw", "lawn"), ("it 's a - dog", "its a-dog")]
    rows = []
    for pred, gt in pairs:
        pl = [charset.index(c) for c in pred]
        gl = [charset.index(c) for c in gt]
        rows.append({"pred": pred, "gt": gt, "cer": ns["character_error_rate"](pred, gt),
                     "clean_pred": ns["process_pred_string"](pred), "clean_gt": ns["process_pred_string"](gt),
                     "pred_words": ns["split_labels_into_words"](pl, charset), "gt_words": ns["split_labels_into_words"](gl, charset),
                     "wer_ref_call": ns["word_error_rate"](ns["split_labels_into_words"](gl, charset),
                                                           ns["split_labels_into_words"](pl, charset))})
    return {"charset": charset, "rows": rows}


def nms_decode_case():
    """evaluation.py:92-160 convert_output_to_pred with args.NMS_inference (top-900 (query, class) pairs, class-agnostic NMS, score
    threshold, sort by cx): the function is extracted by AST and run unmodified with the globals it reads (args, postprocessors,
    dataset_val, box_ops) bound to the reference's own PostProcess / box_ops."""
    import types as _t
    ref_shims.load_reference()
    from models.dino.dino import PostProcess          # reference classes
    from util import box_ops
    src = open(os.path.join(REF, "evaluation.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "convert_output_to_pred"][0]
    ns = {"torch": torch, "box_ops": box_ops, "args": _t.SimpleNamespace(NMS_inference=True),
          "postprocessors": {"bbox": PostProcess()}, "dataset_val": _t.SimpleNamespace(charset=list(range(20)))}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "evaluation.py", "exec"), ns)
    g = torch.Generator().manual_seed(9)
    out = {}
    for i, (TH, NM) in enumerate([(0.3, 0.5), (0.5, 0.2), (0.05, 0.9)]):
        logits = torch.randn(1, 120, 20, generator=g) * 2.0 - 2.0
        cx = torch.rand(1, 120, 1, generator=g)
        boxes = torch.cat((cx, 0.5 + 0.02 * torch.randn(1, 120, 1, generator=g), 0.02 + 0.03 * torch.rand(1, 120, 1, generator=g),
                           0.6 + 0.2 * torch.rand(1, 120, 1, generator=g)), -1)
        preds, labels = ns["convert_output_to_pred"]({"pred_logits": logits, "pred_boxes": boxes}, None, list(range(20)), TH=TH, NM=NM)
        assert preds == labels and len(labels) > 3
        out["nms%d_logits" % i], out["nms%d_boxes" % i] = logits.numpy(), boxes.numpy()
        out["nms%d_labels" % i] = np.array(labels, dtype=np.int64)
        out["nms%d_th_nm" % i] = np.array([TH, NM])
    return out


def main():
    out = transforms_case()
    out.update(ngram_case())
    out.update(nms_decode_case())
    np.savez_compressed(os.path.join(HERE, "io.npz"), **out)
    json.dump(metric_case(), open(os.path.join(HERE, "io_metrics.json"), "w"), indent=1, ensure_ascii=False)
    print("wrote io.npz (%d arrays), io_metrics.json" % len(out))


if __name__ == "__main__":
    main()

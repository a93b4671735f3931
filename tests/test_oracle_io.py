"""CPU: (1) the I/O oracle (oracle/io_ref.py) against the golden vectors the unmodified reference produced
(tests/golden/make_golden_io.py); (2) the product's host-side logic for SURVEY §8f rows 2-4 -- the resize size rule, u8 packing,
width bucketing, CER / WER -- against the same fixtures and the oracle.  No CUDA here."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import io_ref

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


@pytest.fixture(scope="module")
def io(golden_dir):
    return np.load(os.path.join(golden_dir, "io.npz"))


@pytest.fixture(scope="module")
def metrics(golden_dir):
    return json.load(open(os.path.join(golden_dir, "io_metrics.json")))


def test_oracle_totensor_normalize_pad_bit_exact(io):
    ts = [io_ref.to_tensor_normalize(io["img%d_u8" % i], MEAN, STD) for i in range(4)]
    batch, mask = io_ref.nested_batch(ts)
    assert torch.equal(batch, torch.from_numpy(io["batch"])) and torch.equal(mask, torch.from_numpy(io["mask"]))
    ts = [io_ref.to_tensor_normalize(io["rgb%d_u8" % i], MEAN, STD) for i in range(2)]
    batch, mask = io_ref.nested_batch(ts)
    assert torch.equal(batch, torch.from_numpy(io["rgb_batch"])) and torch.equal(mask, torch.from_numpy(io["rgb_mask"]))


def test_size_rule_oracle_and_product(io):
    from dtlr_b200.input import get_size_with_aspect_ratio
    for w, h, size, mx, oh, ow in io["size_table"].tolist():
        assert io_ref.resized_size((w, h), size, mx) == (oh, ow)
        assert get_size_with_aspect_ratio((w, h), size, mx) == (oh, ow)


def test_oracle_ngram_new_pred_logits(io):
    lg, bx = torch.from_numpy(io["np_logits"]), torch.from_numpy(io["np_boxes"])
    for mult in (1, 2):
        ref = torch.from_numpy(io["new_pred_x%d" % mult])
        assert ((ref[..., 0] - 0.003).abs() < 1e-9).any() and ((ref[..., 0] - 0.003).abs() > 1e-4).any()   # both blank branches
        assert torch.allclose(io_ref.new_pred_logits(lg, bx, mult), ref, rtol=1e-6, atol=1e-7)


def test_metrics_oracle_and_product(metrics):
    from dtlr_b200 import evaluation as ev
    cs = metrics["charset"]
    for r in metrics["rows"]:
        pl, gl = [cs.index(c) for c in r["pred"]], [cs.index(c) for c in r["gt"]]
        for mod_cer, mod_clean, mod_split, mod_wer in ((io_ref.cer, io_ref.clean_string, io_ref.split_words, io_ref.wer),
                                                       (ev.character_error_rate, ev.process_pred_string, ev.split_labels_into_words,
                                                        ev.word_error_rate)):
            assert mod_cer(r["pred"], r["gt"]) == r["cer"]
            assert mod_clean(r["pred"]) == r["clean_pred"] and mod_clean(r["gt"]) == r["clean_gt"]
            assert mod_split(pl, cs) == r["pred_words"] and mod_split(gl, cs) == r["gt_words"]
            assert mod_wer(mod_split(gl, cs), mod_split(pl, cs)) == r["wer_ref_call"]
    assert ev.levenshtein_distance("kitten", "sitting") == 3 and ev.levenshtein_distance([], [1, 2]) == 2


def test_bucket_batches_properties():
    from dtlr_b200.evaluation import bucket_batches
    rng = np.random.default_rng(1)
    widths = rng.integers(60, 1400, 257).tolist()
    for bs, mult in ((64, 32), (7, 1), (1, 32), (300, 64)):
        batches = bucket_batches(widths, bs, mult)
        flat = [i for b in batches for i in b]
        assert sorted(flat) == list(range(len(widths)))                         # a partition
        assert all(1 <= len(b) <= bs for b in batches)
        for b in batches:                                                       # one bucket per batch, ascending widths
            ws = [widths[i] for i in b]
            assert ws == sorted(ws)
            assert len({(w + mult - 1) // mult for w in ws}) == 1
        firsts = [widths[b[0]] for b in batches]
        assert firsts == sorted(firsts)
    heights = rng.integers(40, 200, len(widths)).tolist()                      # real evaluation geometry: (W, H) buckets
    batches = bucket_batches(widths, 16, 64, heights=heights, height_multiple=8)
    assert sorted(i for b in batches for i in b) == list(range(len(widths)))
    for b in batches:
        assert 1 <= len(b) <= 16
        assert len({((widths[i] + 63) // 64, (heights[i] + 7) // 8) for i in b}) == 1
    with pytest.raises(ValueError):
        bucket_batches([1, 2], 4, heights=[1])
    assert bucket_batches([], 4) == []
    capped = bucket_batches([100, 101, 127, 128], 8, 128, max_pad_frac=0.1)
    assert [len(b) for b in capped] == [2, 2] or sum(len(b) for b in capped) == 4
    with pytest.raises(ValueError):
        bucket_batches([1], 0)


def test_pack_u8_layout_and_errors(io):
    from dtlr_b200.input import pack_u8
    imgs = [io["img%d_u8" % i] for i in range(4)]
    packed, off, sizes, ch = pack_u8(imgs)
    assert ch == 1 and packed.dtype == np.uint8 and packed.size == sum(im.size for im in imgs)
    for im, o, (h, w) in zip(imgs, off, sizes):
        assert (h, w) == im.shape and np.array_equal(packed[o:o + im.size].reshape(h, w), im)
    packed, off, sizes, ch = pack_u8([io["rgb0_u8"], torch.from_numpy(io["rgb1_u8"])])
    assert ch == 3 and off[1] == io["rgb0_u8"].size
    with pytest.raises(ValueError):
        pack_u8([io["img0_u8"], io["rgb0_u8"]])
    with pytest.raises(TypeError):
        pack_u8([io["img0_u8"].astype(np.float32)])
    with pytest.raises(ValueError):
        pack_u8([])


RESIZE_CASES = [(57, 913, 40, 640), (61, 2011, 31, 1021), (33, 301, 40, 364), (120, 1750, 91, 1333), (40, 512, 40, 512), (30, 100, 60, 250),
                (94, 1333, 94, 1000), (17, 23, 5, 7), (5, 7, 17, 23), (128, 2200, 78, 1340)]


def test_oracle_resize_is_bit_identical_to_pil():
    """third-party arithmetic (Pillow Resample.c) restated in oracle/io_ref.py, pinned against the PIL installed here -- the library
    torchvision F.resize calls for the reference's PIL images (datasets/transforms.py:107-108)."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    for h, w, oh, ow in RESIZE_CASES:
        for shape in ((h, w), (h, w, 3)):
            a = rng.integers(0, 256, shape, dtype=np.uint8)
            ref = np.asarray(Image.fromarray(a).resize((ow, oh), Image.BILINEAR))
            assert np.array_equal(io_ref.pil_resize_bilinear_u8(a, oh, ow), ref), (shape, oh, ow)


def test_product_resample_tables_match_oracle_and_pil():
    """dtlr_b200.input.resample_tables (the host half of dtlr_resize_u8_bilinear) == the oracle's tables, coefficient for
    coefficient, and the two integer passes over those tables (numpy stand-in for the kernels) reproduce PIL."""
    from dtlr_b200 import input as din
    for n_in, n_out in [(913, 640), (2011, 1021), (57, 40), (61, 31), (301, 364), (33, 40), (1750, 1333), (120, 91), (512, 512), (100, 250),
                        (7, 23), (23, 7), (2200, 1340), (1, 5), (5, 1), (1333, 1333)]:
        for a, b in zip(io_ref.resample_coeffs(n_in, n_out), din.resample_tables(n_in, n_out)):
            assert a.dtype == b.dtype and np.array_equal(a, b), (n_in, n_out)
    rng = np.random.default_rng(1)
    for h, w, oh, ow in RESIZE_CASES[:6]:
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(din.apply_tables_u8(a, oh, ow), io_ref.pil_resize_bilinear_u8(a, oh, ow))


def test_resize_plan_layout_drives_a_kernel_emulation_to_the_pil_result():
    """the host/kernel contract of dtlr_resize_u8_bilinear: a literal numpy transcription of resize_pass_kernel (csrc/input.cu), fed
    with GpuResizer.plan()'s packed input / meta / tables for a ragged batch, reproduces the oracle's (== PIL's) images."""
    from dtlr_b200.input import GpuResizer
    rng = np.random.default_rng(4)
    for ch in (1, 3):
        imgs = [rng.integers(0, 256, (h, w) if ch == 1 else (h, w, 3), dtype=np.uint8) for h, w in [(9, 31), (12, 50), (7, 7)]]
        sizes = [(5, 20), (12, 33), (14, 3)]
        packed, meta, tab, tmp_bytes, out_bytes, c, _ = GpuResizer("cpu").plan(imgs, sizes)
        assert c == ch and meta.shape == (3, 12)
        tmp, out = np.zeros(tmp_bytes, np.uint8), np.zeros(out_bytes, np.uint8)
        for m in meta.tolist():
            in_off, tmp_off, out_off, h, w, oh, ow, xoff, yoff, ksx, ksy, _r = m
            for vertical in (False, True):
                src, so = (tmp, tmp_off) if vertical else (packed, in_off)
                dst, do = (out, out_off) if vertical else (tmp, tmp_off)
                n_out, ks, t = (oh, ksy, tab[yoff:]) if vertical else (ow, ksx, tab[xoff:])
                for oy in range(oh if vertical else h):
                    for ox in range(ow):
                        o = oy if vertical else ox
                        lo, cnt = int(t[o]), int(t[n_out + o])
                        for cc in range(ch):
                            acc = 1 << 21
                            for i in range(cnt):
                                p = ((lo + i) * ow + ox) * ch if vertical else (oy * w + lo + i) * ch
                                acc += int(src[so + p + cc]) * int(t[2 * n_out + o * ks + i])
                            dst[do + (oy * ow + ox) * ch + cc] = min(255, max(0, acc >> 22))
        for im, (oh, ow), m in zip(imgs, sizes, meta.tolist()):
            got = out[m[2]:m[2] + oh * ow * ch].reshape((oh, ow) if ch == 1 else (oh, ow, 3))
            assert np.array_equal(got, io_ref.pil_resize_bilinear_u8(im, oh, ow))


def test_nms_decode_oracle_and_product_match_reference(io):
    """evaluation.py:94-115 (NMS_inference branch) as run by the unmodified reference (golden), the oracle restatement, and the
    product's batched `evaluation.nms_decode` over dtlr_b200's PostProcess (torch ops -- here on CPU tensors, on the GPU in use)."""
    from dtlr_b200 import dino, evaluation
    for i in range(3):
        lg, bx = torch.from_numpy(io["nms%d_logits" % i]), torch.from_numpy(io["nms%d_boxes" % i])
        th, nm = io["nms%d_th_nm" % i].tolist()
        want = io["nms%d_labels" % i].tolist()
        assert io_ref.nms_decode(lg, bx, th, nm) == want
        got = evaluation.nms_decode({"pred_logits": lg, "pred_boxes": bx}, dino.PostProcess(), th, nm)
        assert got == [want]
    # batched: two different images in one call
    lg = torch.cat([torch.from_numpy(io["nms%d_logits" % i]) for i in (0, 1)])
    bx = torch.cat([torch.from_numpy(io["nms%d_boxes" % i]) for i in (0, 1)])
    th, nm = io["nms0_th_nm"].tolist()
    got = evaluation.nms_decode({"pred_logits": lg, "pred_boxes": bx}, dino.PostProcess(), th, nm)
    assert got[0] == io["nms0_labels"].tolist() and got[1] == io_ref.nms_decode(lg[1:], bx[1:], th, nm)

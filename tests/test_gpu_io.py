"""GPU: the stages either side of the forward through the C ABI against the reference's golden vectors
(tests/golden/io.npz, produced by the unmodified reference) and the CPU oracle (oracle/io_ref.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import io_ref

pytestmark = pytest.mark.gpu
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


@pytest.fixture(scope="module")
def io(golden_dir):
    return np.load(os.path.join(golden_dir, "io.npz"))


def test_gpu_preprocess_bit_identical_to_reference_chain(io):
    """dtlr_preprocess_u8 == ToTensor + Normalize + nested_tensor_from_tensor_list of the reference, bit for bit."""
    from dtlr_b200.input import GpuPreprocessor
    prep = GpuPreprocessor("cuda")
    nt = prep([io["img%d_u8" % i] for i in range(4)])
    assert nt.tensors.dtype == torch.float32 and nt.mask.dtype == torch.bool and not nt.nopad
    assert torch.equal(nt.tensors.cpu(), torch.from_numpy(io["batch"])) and torch.equal(nt.mask.cpu(), torch.from_numpy(io["mask"]))
    assert prep.h2d_bytes == sum(io["img%d_u8" % i].size for i in range(4)) + 24 * 4      # 1 byte per pixel crosses PCIe
    nt = prep([io["rgb0_u8"], io["rgb1_u8"]])                                             # interleaved RGB, reused staging
    assert torch.equal(nt.tensors.cpu(), torch.from_numpy(io["rgb_batch"])) and torch.equal(nt.mask.cpu(), torch.from_numpy(io["rgb_mask"]))


def test_gpu_preprocess_width_rounding_full_size_and_all_u8_values():
    from dtlr_b200.input import GpuPreprocessor
    rng = np.random.default_rng(2)
    imgs = [rng.integers(0, 256, (40, w), dtype=np.uint8) for w in (1000, 1024, 37, 1)]
    imgs[2][0, :] = np.arange(37) * 7 % 256
    imgs.append(np.arange(256, dtype=np.uint8).reshape(1, 256).repeat(3, 0))            # every u8 value, a short image
    nt = GpuPreprocessor("cuda", pad_w_multiple=32)(imgs)
    ref, mask = io_ref.nested_batch([io_ref.to_tensor_normalize(im, MEAN, STD) for im in imgs], pad_to_w=1024)
    assert nt.tensors.shape == (5, 3, 40, 1024)
    assert torch.equal(nt.tensors.cpu(), ref) and torch.equal(nt.mask.cpu(), mask)
    # odd batch width (scalar store path), height rounding, and the side-stream variant (caller's stream waits for it)
    side = torch.cuda.Stream()
    odd = [imgs[2], imgs[3], imgs[4][:, :33]]
    for stream in (None, side):
        nt = GpuPreprocessor("cuda", pad_h_multiple=16)(odd, stream=stream)
        ref, mask = io_ref.nested_batch([io_ref.to_tensor_normalize(im, MEAN, STD) for im in odd], pad_to_h=48)
        assert nt.tensors.shape == (3, 3, 48, 37)
        assert torch.equal(nt.tensors.cpu(), ref) and torch.equal(nt.mask.cpu(), mask)
    same = GpuPreprocessor("cuda")([imgs[1], imgs[1]])                                   # equal sizes -> dense-batch hint
    assert same.nopad and not bool(same.mask.any())


def test_gpu_preprocess_rejects_bad_input():
    from dtlr_b200 import _lib, ops
    with pytest.raises(_lib.DtlrError):
        ops.preprocess_u8(torch.zeros(4, dtype=torch.uint8), torch.zeros(1, dtype=torch.int64), torch.ones(1, 2, dtype=torch.int32),
                          1, 1, 1, 1, MEAN, STD)                                          # CPU tensors
    z = torch.zeros(4, dtype=torch.uint8, device="cuda")
    with pytest.raises(_lib.DtlrError):
        ops.preprocess_u8(z, torch.zeros(1, dtype=torch.int64, device="cuda"), torch.ones(1, 2, dtype=torch.int32, device="cuda"),
                          2, 1, 1, 1, MEAN, STD)                                          # 2 channels


def test_ngram_feed_matches_reference_helper(io):
    """dtlr_b200.ngram.get_new_pred_logits == reference ngram/prediction_helpers.py:get_new_pred_logits (golden), layout (B,Q,C+1)."""
    from dtlr_b200 import ngram, ops
    out = {"pred_logits": torch.from_numpy(io["np_logits"]).cuda(), "pred_boxes": torch.from_numpy(io["np_boxes"]).cuda()}
    for mult in (1, 2):
        ref = torch.from_numpy(io["new_pred_x%d" % mult])
        got = ngram.get_new_pred_logits(out, mult).cpu()
        assert got.shape == ref.shape and torch.allclose(got, ref, rtol=1e-5, atol=1e-6)
        frames = ops.ctc_decode(out["pred_logits"], out["pred_boxes"], 0.003, prob_scale=float(mult)).cpu()
        assert torch.equal(frames.long(), ref.argmax(-1))


def test_bucketed_evaluator_equals_direct_batches_and_oracle_metrics():
    """LineEvaluator (width buckets, GPU input stage, overlapped frame download, eps = 0.03/C) returns for every line exactly what a
    direct call on that line's batch returns, in input order; CER / WER equal the oracle's on those predictions."""
    from gpu_common import build_model
    from dtlr_b200 import evaluation, ops
    from dtlr_b200.misc import NestedTensor
    model, _, _ = build_model(100)
    model.eval()
    with torch.no_grad():                        # class bias shifted so that blank and characters both occur
        for ce in model.class_embed:
            ce.bias.add_(3.5)
    rng = np.random.default_rng(5)
    widths = [704, 333, 350, 352, 97, 699, 640, 351, 100, 705]
    images = [rng.integers(0, 256, (40 if i % 3 else 37, w), dtype=np.uint8) for i, w in enumerate(widths)]
    C = 166
    charset = [chr(0x21 + i) for i in range(C)]
    charset[5] = " "
    gts = [rng.integers(0, C, rng.integers(0, 30)).tolist() for _ in images]
    for graph in (False, True):
        model.use_cuda_graph = graph
        ev = evaluation.LineEvaluator(model, charset, batch_size=3, width_multiple=32)
        res = ev.evaluate(images, gts)
        preds = res["preds"]
        assert len(preds) == len(images) and all(p is not None for p in preds)
        seen = 0
        batches = ev.batches(images)
        assert batches == evaluation.bucket_batches(widths, 3, 32, heights=[im.shape[0] for im in images], height_multiple=8)
        for idx in batches:
            ts = [io_ref.to_tensor_normalize(images[i], MEAN, STD) for i in idx]
            wpad = (max(widths[i] for i in idx) + 31) // 32 * 32
            hpad = (max(images[i].shape[0] for i in idx) + 7) // 8 * 8
            x, m = io_ref.nested_batch(ts, pad_to_w=wpad, pad_to_h=hpad)
            with torch.no_grad():
                out = model(NestedTensor(x.cuda(), m.cuda()))
                frames = ops.ctc_decode(out["pred_logits"], out["pred_boxes"], 0.03 / C).cpu()
            for row, i in zip(frames.tolist(), idx):
                assert preds[i] == [v - 1 for v in row if v != 0]
                seen += 1
        assert seen == len(images)
        assert any(len(p) > 0 for p in preds)
        # metrics with the oracle's definitions on the same predictions
        d = l = 0
        wers = []
        for p, g in zip(preds, gts):
            ps, gs = "".join(charset[c] for c in p), "".join(charset[c] for c in g)
            d += io_ref.edit_distance(io_ref.clean_string(gs), io_ref.clean_string(ps))
            l += len(io_ref.clean_string(gs))
            wers.append(io_ref.wer(io_ref.split_words(g, charset), io_ref.split_words(p, charset)))
        assert res["cer"] == d / max(l, 1) and abs(res["wer"] - sum(wers) / len(wers)) < 1e-12


def test_gpu_resize_bit_identical_to_pil_restatement():
    """dtlr_resize_u8_bilinear (ragged batch, 1 and 3 channels, up- and down-scaling) == the oracle's restatement of PIL's
    Image.resize(BILINEAR) bit for bit, and GpuPreprocessor.resized == the reference's whole evaluation transform."""
    from dtlr_b200.input import GpuPreprocessor, GpuResizer, get_size_with_aspect_ratio
    rng = np.random.default_rng(3)
    for ch in (1, 3):
        imgs = [rng.integers(0, 256, (h, w) if ch == 1 else (h, w, 3), dtype=np.uint8) for h, w in [(57, 913), (61, 2011), (33, 301), (120, 1750), (17, 23)]]
        sizes = [(40, 640), (31, 1021), (40, 364), (91, 1333), (5, 7)]
        d_out, d_off, hw, c = GpuResizer("cuda")(imgs, sizes)
        flat, offs = d_out.cpu().numpy(), d_off.cpu().tolist()
        assert c == ch and hw.cpu().tolist() == [list(s) for s in sizes]
        for im, (oh, ow), o in zip(imgs, sizes, offs):
            got = flat[o:o + oh * ow * ch].reshape((oh, ow) if ch == 1 else (oh, ow, 3))
            assert np.array_equal(got, io_ref.pil_resize_bilinear_u8(im, oh, ow))
    # the whole evaluation transform: resize(800, max 1333) + ToTensor + Normalize + pad == the oracle chain
    scans = [rng.integers(0, 256, (h, w), dtype=np.uint8) for h, w in [(120, 1750), (90, 1500), (140, 2100)]]
    nt = GpuPreprocessor("cuda").resized(scans, 800, 1333)
    ts = []
    for a in scans:
        oh, ow = get_size_with_aspect_ratio((a.shape[1], a.shape[0]), 800, 1333)
        ts.append(io_ref.to_tensor_normalize(io_ref.pil_resize_bilinear_u8(a, oh, ow), MEAN, STD))
    ref, mask = io_ref.nested_batch(ts)
    assert torch.equal(nt.tensors.cpu(), ref) and torch.equal(nt.mask.cpu(), mask)

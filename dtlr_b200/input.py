"""GPU input stage: resized 8-bit line images -> the NestedTensor DINO.forward consumes.

The reference prepares every image on the host (datasets/IAM.py:85-112 + datasets/transforms.py): PIL `convert("RGB")`,
`RandomResize([max(scales)], max_size)`, `ToTensor`, `Normalize(IMAGENET)`, then `collate_fn` pads the batch
(util/misc.py:285-289, 375-397) and 12 bytes per pixel cross PCIe.  Here the host only decodes and resizes (PIL's fixed-point
antialiasing filter is not restated); the u8 pixels are packed into ONE pinned buffer, cross PCIe at 1 byte per pixel for
grayscale lines, and `dtlr_preprocess_u8` writes the normalised, padded fp32 batch and its mask -- bit-identical to the
reference chain on the same u8 pixels.
"""
import numpy as np
import torch

from . import ops
from .misc import NestedTensor

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # datasets/IAM.py:110-112
IMAGENET_STD = (0.229, 0.224, 0.225)


def get_size_with_aspect_ratio(image_size, size, max_size=None):
    """reference datasets/transforms.py:81-99: (w, h) of the PIL image -> (oh, ow) with the short side = `size`, the long side
    capped at `max_size`."""
    w, h = image_size
    if max_size is not None:
        min_original_size = float(min((w, h)))
        max_original_size = float(max((w, h)))
        if max_original_size / min_original_size * size > max_size:
            size = int(round(max_size * min_original_size / max_original_size))
    if (w <= h and w == size) or (h <= w and h == size):
        return (h, w)
    if w < h:
        ow = size
        oh = int(size * h / w)
    else:
        oh = size
        ow = int(size * w / h)
    return (oh, ow)


def _as_u8_array(img):
    a = img.numpy() if torch.is_tensor(img) else np.asarray(img)
    if a.dtype != np.uint8:
        raise TypeError("GpuPreprocessor takes 8-bit images (got %s)" % a.dtype)
    if a.ndim == 2:
        a = a[:, :, None]
    if a.ndim != 3 or a.shape[2] not in (1, 3):
        raise ValueError("image must be (H,W), (H,W,1) or (H,W,3); got %s" % (a.shape,))
    return np.ascontiguousarray(a)


def pack_u8(images):
    """list of (H,W[,1|3]) u8 arrays / tensors / PIL images -> (packed 1-D u8 array, int64 offsets, int32 (B,2) sizes, channels).
    All images must have the same channel count."""
    arrs = [_as_u8_array(im) for im in images]
    if not arrs:
        raise ValueError("empty batch")
    ch = arrs[0].shape[2]
    if any(a.shape[2] != ch for a in arrs):
        raise ValueError("all images of a batch must have the same number of channels")
    sizes = np.array([a.shape[:2] for a in arrs], dtype=np.int32)
    nbytes = np.array([a.size for a in arrs], dtype=np.int64)
    offsets = np.concatenate(([0], np.cumsum(nbytes)[:-1])).astype(np.int64)
    packed = np.concatenate([a.reshape(-1) for a in arrs])
    return packed, offsets, sizes, ch


PRECISION_BITS = 32 - 8 - 2          # Pillow Resample.c: 8-bit samples, 22-bit fixed-point coefficients


def resample_tables(in_size, out_size):
    """coefficient table of one axis of PIL's bilinear (triangle-filter, antialiased) resize, as Pillow's precompute_coeffs +
    normalize_coeffs_8bpc build it (double precision, same operation order; the reference resizes with torchvision F.resize on PIL
    images, datasets/transforms.py:107-108).  Returns (xmin[out] int32, count[out] int32, coef[out, ksize] int32)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum(np.trunc(center - support + 0.5).astype(np.int64), 0)       # C (int) casts truncate toward zero
    xmax = np.minimum(np.trunc(center + support + 0.5).astype(np.int64), in_size)
    cnt = xmax - xmin
    x = np.arange(ksize, dtype=np.int64)[None, :]
    t = np.abs(((x + xmin[:, None]).astype(np.float64) - center[:, None] + 0.5) * ss)
    w = np.where(t < 1.0, 1.0 - t, 0.0)
    w = np.where(x < cnt[:, None], w, 0.0)
    ww = np.zeros(out_size, dtype=np.float64)
    for k in range(ksize):                                                          # sequential sum, as the C loop
        ww = ww + w[:, k]
    kk = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    coef = np.trunc(0.5 + kk * float(1 << PRECISION_BITS)).astype(np.int32)         # bilinear weights are >= 0
    coef = np.where(x < cnt[:, None], coef, 0).astype(np.int32)
    return xmin.astype(np.int32), cnt.astype(np.int32), coef


def apply_tables_u8(img, out_h, out_w):
    """host check of the tables only (numpy; used by the CPU tests, never by the product path): the two integer passes the
    CUDA kernels run, on an (H,W[,C]) u8 array."""
    a = np.asarray(img)
    squeeze = a.ndim == 2
    if squeeze:
        a = a[:, :, None]
    for axis, n_out in ((1, out_w), (0, out_h)):
        xmin, cnt, coef = resample_tables(a.shape[axis], n_out)
        src = np.moveaxis(a, axis, 0).astype(np.int64)
        idx = np.minimum(xmin[:, None] + np.arange(coef.shape[1])[None, :], src.shape[0] - 1)
        acc = (1 << (PRECISION_BITS - 1)) + np.einsum("ok,ok...->o...", coef.astype(np.int64), src[idx])
        a = np.moveaxis(np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8), 0, axis)
    return a[:, :, 0] if squeeze else a


class GpuResizer:
    """ragged batch of u8 images -> PIL-exact bilinear resize on the GPU (dtlr_resize_u8_bilinear).  Returns the packed resized
    images as device tensors in the layout dtlr_preprocess_u8 takes: (packed u8, int64 offsets, int32 (B,2) sizes, channels)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self._tables = {}                      # (in_size, out_size) -> flat int32 table, cached across batches

    def _table(self, n_in, n_out):
        key = (int(n_in), int(n_out))
        t = self._tables.get(key)
        if t is None:
            xmin, cnt, coef = resample_tables(*key)
            t = (np.concatenate((xmin, cnt, coef.reshape(-1))).astype(np.int32), coef.shape[1])
            if len(self._tables) > 4096:
                self._tables.clear()
            self._tables[key] = t
        return t

    def plan(self, images, out_sizes):
        """host half: (packed u8 input, int64 meta [B,12], flat int32 tables, tmp bytes, out bytes, channels, input sizes) in the
        layout include/dtlr_b200.h documents for dtlr_resize_u8_bilinear."""
        packed, in_off, sizes, ch = pack_u8(images)
        B = len(sizes)
        meta = np.zeros((B, 12), dtype=np.int64)
        tabs, tab_off, tmp_off, out_off = [], 0, 0, 0
        for b in range(B):
            h, w = int(sizes[b, 0]), int(sizes[b, 1])
            oh, ow = int(out_sizes[b][0]), int(out_sizes[b][1])
            tx, ksx = self._table(w, ow)
            ty, ksy = self._table(h, oh)
            meta[b] = (in_off[b], tmp_off, out_off, h, w, oh, ow, tab_off, tab_off + tx.size, ksx, ksy, 0)
            tabs += [tx, ty]
            tab_off += tx.size + ty.size
            tmp_off += h * ow * ch
            out_off += oh * ow * ch
        return packed, meta, np.concatenate(tabs), tmp_off, out_off, ch, sizes

    def __call__(self, images, out_sizes):
        """images: list of (H,W[,1|3]) u8; out_sizes: list of (oh, ow)."""
        packed, meta, tab, tmp_off, out_off, ch, sizes = self.plan(images, out_sizes)
        B = len(sizes)
        dev = self.device
        d_in = torch.from_numpy(packed).to(dev, non_blocking=True)
        d_meta = torch.from_numpy(meta).to(dev, non_blocking=True)
        d_tab = torch.from_numpy(tab).to(dev, non_blocking=True)
        d_tmp = torch.empty(max(tmp_off, 1), dtype=torch.uint8, device=dev)
        d_out = torch.empty(max(out_off, 1), dtype=torch.uint8, device=dev)
        ops.resize_u8_bilinear(d_in, d_meta, d_tab, d_tmp, d_out, B, ch, int(sizes[:, 0].max()),
                               max(int(s[0]) for s in out_sizes), max(int(s[1]) for s in out_sizes))
        hw = torch.tensor([[int(s[0]), int(s[1])] for s in out_sizes], dtype=torch.int32).to(dev, non_blocking=True)
        return d_out, d_meta[:, 2].contiguous(), hw, ch


class GpuPreprocessor:
    """host u8 images -> NestedTensor on `device` (one pinned staging buffer, one H2D copy, one kernel).
    `pad_w_multiple` / `pad_h_multiple` round the batch size up (fewer distinct shapes -> fewer captured CUDA graphs); the extra
    columns / rows are ordinary padding (mask True), exactly what a larger image in the batch would cause in the reference.
    With `stream=` the copy and the kernel run on that side stream (overlapping whatever the caller's stream is computing) and the
    caller's current stream is made to wait for them."""

    def __init__(self, device, mean=IMAGENET_MEAN, std=IMAGENET_STD, pad_w_multiple=1, pad_h_multiple=1):
        self.device = torch.device(device)
        self.mean, self.std = tuple(mean), tuple(std)
        self.pad_w_multiple = int(pad_w_multiple)
        self.pad_h_multiple = int(pad_h_multiple)
        self._stage = None
        self._meta = None
        self._copied = None             # event after the last H2D copies: the pinned staging buffers may be rewritten after it
        self.h2d_bytes = 0

    def _staging(self, nbytes, B):
        if self._stage is None or self._stage.numel() < nbytes:
            self._stage = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8).pin_memory()
        if self._meta is None or self._meta.numel() < 3 * B:
            self._meta = torch.empty(max(3 * B, 384), dtype=torch.int64).pin_memory()
        return self._stage, self._meta

    def resized(self, images, size, max_size=None, stream=None):
        """the reference's evaluation transform on the GPU: RandomResize([size], max_size) (datasets/transforms.py:78-108, PIL-exact
        bilinear) + ToTensor + Normalize + padding.  `images` are the ORIGINAL u8 scans."""
        if not hasattr(self, "_resizer"):
            self._resizer = GpuResizer(self.device)
        arrs = [_as_u8_array(im) for im in images]
        out_sizes = [get_size_with_aspect_ratio((a.shape[1], a.shape[0]), size, max_size) for a in arrs]
        consumer = torch.cuda.current_stream(self.device)
        work = stream if stream is not None else consumer
        B = len(arrs)
        mh, mw = self.pad_h_multiple, self.pad_w_multiple
        Hpad = (max(s[0] for s in out_sizes) + mh - 1) // mh * mh
        Wpad = (max(s[1] for s in out_sizes) + mw - 1) // mw * mw
        with torch.cuda.stream(work):
            d_packed, d_off, hw, ch = self._resizer(arrs, out_sizes)
            out, mask = ops.preprocess_u8(d_packed, d_off, hw, ch, B, Hpad, Wpad, self.mean, self.std)
            if stream is not None:
                done = torch.cuda.Event()
                done.record(work)
        if stream is not None:
            consumer.wait_event(done)
            out.record_stream(consumer)
            mask.record_stream(consumer)
        self.h2d_bytes = sum(a.size for a in arrs)
        same = all(s[0] == Hpad and s[1] == Wpad for s in out_sizes)
        return NestedTensor(out, mask, nopad=same)

    def __call__(self, images, stream=None):
        packed, offsets, sizes, ch = pack_u8(images)
        B = len(sizes)
        mh, mw = self.pad_h_multiple, self.pad_w_multiple
        Hpad = (int(sizes[:, 0].max()) + mh - 1) // mh * mh
        Wpad = (int(sizes[:, 1].max()) + mw - 1) // mw * mw
        if self._copied is not None:
            self._copied.synchronize()
        stage, meta = self._staging(packed.size, B)
        stage[:packed.size].copy_(torch.from_numpy(packed))
        meta[:3 * B].copy_(torch.from_numpy(np.concatenate((offsets, sizes[:, 0].astype(np.int64), sizes[:, 1].astype(np.int64)))))
        consumer = torch.cuda.current_stream(self.device)
        work = stream if stream is not None else consumer
        with torch.cuda.stream(work):
            d_packed = stage[:packed.size].to(self.device, non_blocking=True)       # ONE u8 copy + one 24*B-byte copy
            d_meta = meta[:3 * B].to(self.device, non_blocking=True)
            self._copied = torch.cuda.Event()
            self._copied.record(work)
            hw = torch.stack((d_meta[B:2 * B], d_meta[2 * B:]), 1).to(torch.int32).contiguous()
            out, mask = ops.preprocess_u8(d_packed, d_meta[:B].contiguous(), hw, ch, B, Hpad, Wpad, self.mean, self.std)
            if stream is not None:
                done = torch.cuda.Event()
                done.record(work)
        if stream is not None:
            consumer.wait_event(done)
            out.record_stream(consumer)
            mask.record_stream(consumer)
        self.h2d_bytes = packed.size + 24 * B
        same = bool((sizes[:, 0] == Hpad).all() and (sizes[:, 1] == Wpad).all())
        return NestedTensor(out, mask, nopad=same)

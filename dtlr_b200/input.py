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

"""Host-side helpers the model boundary needs (mirrors of reference util/misc.py for the hot path only)."""
from typing import List, Optional

import torch
from torch import Tensor


class NestedTensor(object):
    """mirror of reference util/misc.py:301-372 (tensors + padding mask, True on padding)."""

    def __init__(self, tensors, mask: Optional[Tensor], nopad: bool = False):
        self.tensors = tensors
        self.mask = mask
        self.nopad = nopad      # dtlr_b200 hint: the mask is known (on the host) to be all-False -> padding passes are skipped

    def to(self, device):
        m = self.mask.to(device) if self.mask is not None else None
        return NestedTensor(self.tensors.to(device), m, self.nopad)

    def decompose(self):
        return self.tensors, self.mask

    @property
    def device(self):
        return self.tensors.device

    def __repr__(self):
        return str(self.tensors)


def nested_tensor_from_tensor_list(tensor_list: List[Tensor]):
    """reference util/misc.py:375-397: zero-pad every (C,h,w) image to the batch max, mask True on padding.
    (a (B,C,H,W) tensor iterates as a list of equally sized images -> all-False mask)."""
    if torch.is_tensor(tensor_list) and tensor_list.dim() == 4:
        b, c, h, w = tensor_list.shape
        return NestedTensor(tensor_list, torch.zeros((b, h, w), dtype=torch.bool, device=tensor_list.device), nopad=True)
    if tensor_list[0].ndim != 3:
        raise ValueError("not supported")
    c = tensor_list[0].shape[0]
    h = max(int(t.shape[1]) for t in tensor_list)
    w = max(int(t.shape[2]) for t in tensor_list)
    dtype, device = tensor_list[0].dtype, tensor_list[0].device
    tensor = torch.zeros((len(tensor_list), c, h, w), dtype=dtype, device=device)
    mask = torch.ones((len(tensor_list), h, w), dtype=torch.bool, device=device)
    for img, pad_img, m in zip(tensor_list, tensor, mask):
        pad_img[: img.shape[0], : img.shape[1], : img.shape[2]].copy_(img)
        m[: img.shape[1], : img.shape[2]] = False
    same = all(int(t.shape[1]) == h and int(t.shape[2]) == w for t in tensor_list)
    return NestedTensor(tensor, mask, nopad=same)


def collate_fn(batch):
    """reference util/misc.py:285-289: DataLoader collate -- (images, targets) pairs -> (NestedTensor, tuple of targets)."""
    images, *rest = zip(*batch)
    return (nested_tensor_from_tensor_list(list(images)), *rest)


def inverse_sigmoid(x, eps=1e-3):
    """reference util/misc.py:575-579"""
    x = x.clamp(min=0, max=1)
    x1 = x.clamp(min=eps)
    x2 = (1 - x).clamp(min=eps)
    return torch.log(x1 / x2)

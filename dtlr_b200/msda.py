"""Multi-scale deformable attention: the operator boundary.

 * `ms_deform_attn_forward` / `ms_deform_attn_backward` have the exact signature of the reference's pybind module
   (models/dino/ops/src/vision.cpp:13-16) and `install_as_reference_extension()` registers this module as
   sys.modules['MultiScaleDeformableAttention'], so the reference's own
   models/dino/ops/functions/ms_deform_attn_func.py:18 imports it unchanged.
 * `MSDeformAttnFunction` mirrors reference ops/functions/ms_deform_attn_func.py:21-38.
Everything goes through the C ABI (include/dtlr_b200.h); there is no PyTorch fallback.
"""
import sys
import weakref

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib as L

_shape_cache = {}


def _host_levels(spatial_shapes, level_start_index):
    """The reference kernel reads the two int64 tensors on the device; our launcher wants them on the host (it sizes
    shared memory from them).  One D2H copy per distinct tensor version, cached (the reference itself syncs on
    spatial_shapes on every call, ops/modules/ms_deform_attn.py:92)."""
    key = (id(spatial_shapes), id(level_start_index))
    hit = _shape_cache.get(key)
    if hit is not None:
        r0, r1, v0, v1, payload = hit
        if r0() is spatial_shapes and r1() is level_start_index and v0 == spatial_shapes._version \
                and v1 == level_start_index._version:
            return payload
    if len(_shape_cache) > 64:
        _shape_cache.clear()
    sh = [int(v) for v in spatial_shapes.detach().cpu().reshape(-1).tolist()]
    ls = [int(v) for v in level_start_index.detach().cpu().reshape(-1).tolist()]
    payload = (L.i64_host(sh), L.i64_host(ls), len(ls))
    _shape_cache[key] = (weakref.ref(spatial_shapes), weakref.ref(level_start_index), spatial_shapes._version,
                         level_start_index._version, payload)
    return payload


def msda_forward_raw(value, shapes_host, lsi_host, n_levels, loc, attn, out=None):
    """value (B,S,M,D) cuda contiguous; shapes_host/lsi_host ctypes int64 arrays; loc (B,Lq,M,L,P,2); attn (B,Lq,M,L,P)."""
    L.require_cuda(value, loc, attn)
    L.set_flavor(value.dtype)
    if not (value.is_contiguous() and loc.is_contiguous() and attn.is_contiguous()):
        raise L.DtlrError("ms_deform_attn_forward: value, sampling_loc and attn_weight must be contiguous")
    B, S, M, D = value.shape
    Lq, P = loc.shape[1], loc.shape[4]
    pdt = torch.float64 if value.dtype == torch.float64 else torch.float32
    if loc.dtype != pdt or attn.dtype != pdt:
        raise L.DtlrError("ms_deform_attn_forward: sampling_loc/attn_weight must be %s for %s values" % (pdt, value.dtype))
    if out is None:
        out = torch.empty((B, Lq, M * D), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device):
        rc = L.lib().dtlr_msda_forward(L.ptr(value), shapes_host, lsi_host, L.ptr(loc), L.ptr(attn), L.ptr(out),
                                       B, S, M, D, n_levels, Lq, P, L.dtype_code(value), L.stream_ptr(value.device))
    L.check(rc, "dtlr_msda_forward")
    return out


def msda_forward_fused(value, shapes_host, lsi_host, n_levels, proj, ref, valid_ratios, Lq, P, out=None):
    """MSDA core with the prologue fused (include/dtlr_b200.h: dtlr_msda_forward_fused).  value (B,S,M,32);
    proj fp32 (B*Lq, ld) = offsets | logits; ref fp32 (B*Lq, 2|4); valid_ratios fp32 (B,L,2)."""
    L.require_cuda(value, proj, ref, valid_ratios)
    L.set_flavor(value.dtype)
    B, S, M, D = value.shape
    # value may be a column block of a wider matrix (several layers' value projections from one GEMM): pixel pitch = stride(1)
    assert value.stride(3) == 1 and value.stride(2) == D and value.stride(0) == S * value.stride(1), "unsupported value layout"
    assert proj.dtype in (torch.float32, torch.bfloat16, torch.float16) and ref.dtype == torch.float32 and valid_ratios.dtype == torch.float32
    assert ref.is_contiguous() and valid_ratios.is_contiguous() and proj.stride(1) == 1
    if out is None:
        out = torch.empty((B, Lq, M * D), dtype=value.dtype, device=value.device)
    if L.TIMER is not None:     # bench.py: algorithmic bytes of this call (values + projection rows + reference points + output)
        L.TIMER("msda", float(B * (S * M * D * value.element_size() + Lq * proj.shape[1] * proj.element_size()
                                   + Lq * ref.shape[-1] * 4 + Lq * M * D * value.element_size()) + 96), value.device, True)
    with torch.cuda.device(value.device):
        rc = L.lib().dtlr_msda_forward_fused(L.ptr(value), value.stride(1), shapes_host, lsi_host, L.ptr(proj), proj.stride(0), L.dtype_code(proj), L.ptr(ref),
                                             ref.shape[-1], L.ptr(valid_ratios), L.ptr(out), B, S, M, D, n_levels, Lq, P,
                                             L.dtype_code(value), L.stream_ptr(value.device))
    L.check(rc, "dtlr_msda_forward_fused")
    if L.TIMER is not None:
        L.TIMER("msda", 0.0, value.device, False)
    return out


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step=64):
    """Drop-in for MultiScaleDeformableAttention.ms_deform_attn_forward (reference src/ms_deform_attn.h:21-40).
    im2col_step is accepted and ignored (no batch chunking on B200)."""
    sh, ls, n = _host_levels(spatial_shapes, level_start_index)
    return msda_forward_raw(value, sh, ls, n, sampling_loc, attn_weight)


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step=64):
    """Drop-in for MultiScaleDeformableAttention.ms_deform_attn_backward (reference src/ms_deform_attn.h:42-60)."""
    L.require_cuda(value, sampling_loc, attn_weight, grad_output)
    sh, ls, n = _host_levels(spatial_shapes, level_start_index)
    B, S, M, D = value.shape
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    grad_output = grad_output.contiguous()
    gv = torch.empty_like(value)
    gl = torch.empty_like(sampling_loc)
    ga = torch.empty_like(attn_weight)
    with torch.cuda.device(value.device):
        rc = L.lib().dtlr_msda_backward(L.ptr(value), sh, ls, L.ptr(sampling_loc), L.ptr(attn_weight), L.ptr(grad_output),
                                        L.ptr(gv), L.ptr(gl), L.ptr(ga), B, S, M, D, n, Lq, P, L.dtype_code(value),
                                        L.stream_ptr(value.device))
    L.check(rc, "dtlr_msda_backward")
    return gv, gl, ga


class MSDeformAttnFunction(Function):
    """mirror of reference ops/functions/ms_deform_attn_func.py:21-38"""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step=64):
        ctx.im2col_step = im2col_step
        out = ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                     attention_weights, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        gv, gl, ga = ms_deform_attn_backward(value, shapes, lsi, loc, attn, grad_output, ctx.im2col_step)
        return gv, None, None, gl, ga, None


def install_as_reference_extension():
    """Register this module under the name the reference imports (ops/functions/ms_deform_attn_func.py:18)."""
    sys.modules["MultiScaleDeformableAttention"] = sys.modules[__name__]

"""ORACLE (test infrastructure): ctypes loader for the C restatement of the MSDA core (oracle/msda_ref.c)."""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libmsda_ref.so"])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libmsda_ref.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(value, shapes, lsi, loc, w):
    dt = np.float64 if value.dtype == torch.float64 else np.float32
    v = np.ascontiguousarray(value.detach().cpu().numpy().astype(dt))
    lo = np.ascontiguousarray(loc.detach().cpu().numpy().astype(dt))
    ww = np.ascontiguousarray(w.detach().cpu().numpy().astype(dt))
    sh = np.ascontiguousarray(torch.as_tensor(shapes).cpu().numpy().astype(np.int64))
    ls = np.ascontiguousarray(torch.as_tensor(lsi).cpu().numpy().astype(np.int64))
    B, S, M, D = v.shape
    _, Lq, _, L, P, _ = lo.shape
    return dt, v, sh, ls, lo, ww, (B, S, M, D, L, Lq, P)


_REF_CUDA = None


def _ref_cuda():
    """the reference's OWN CUDA op recompiled for sm_100a (oracle/_ref/, oracle/build_ref_cuda.py): the GPU-side baseline"""
    global _REF_CUDA
    if _REF_CUDA is None:
        from . import build_ref_cuda
        _REF_CUDA = build_ref_cuda.load()
        if _REF_CUDA is None:
            raise RuntimeError("oracle/_ref/MultiScaleDeformableAttention.so not built (python oracle/build_ref_cuda.py)")
    return _REF_CUDA


def msda_forward(value, shapes, lsi, loc, w):
    """value (B,S,M,D), shapes (L,2), lsi (L,), loc (B,Lq,M,L,P,2), w (B,Lq,M,L,P) -> (B,Lq,M*D) torch CPU.
    CUDA tensors go to the reference's own CUDA kernel (bench.py `gpu_reference` leg: reference-shaped torch path on the same GPU)."""
    if value.is_cuda:
        B = value.shape[0]
        return _ref_cuda().ms_deform_attn_forward(value.contiguous(), shapes.to(value.device), lsi.to(value.device),
                                                  loc.contiguous(), w.contiguous(), 64 if B % 64 == 0 else B)
    dt, v, sh, ls, lo, ww, (B, S, M, D, L, Lq, P) = _prep(value, shapes, lsi, loc, w)
    out = np.empty((B, Lq, M * D), dtype=dt)
    fn = _lib().msda_ref_fwd_f64 if dt == np.float64 else _lib().msda_ref_fwd_f32
    fn(_ptr(v), _ptr(sh), _ptr(ls), _ptr(lo), _ptr(ww), _ptr(out), B, S, M, D, L, Lq, P)
    return torch.from_numpy(out)


def msda_backward(value, shapes, lsi, loc, w, grad_out):
    dt, v, sh, ls, lo, ww, (B, S, M, D, L, Lq, P) = _prep(value, shapes, lsi, loc, w)
    go = np.ascontiguousarray(grad_out.detach().cpu().numpy().astype(dt))
    gv = np.zeros_like(v)
    gl = np.zeros_like(lo)
    gw = np.zeros_like(ww)
    fn = _lib().msda_ref_bwd_f64 if dt == np.float64 else _lib().msda_ref_bwd_f32
    fn(_ptr(v), _ptr(sh), _ptr(ls), _ptr(lo), _ptr(ww), _ptr(go), _ptr(gv), _ptr(gl), _ptr(gw),
       B, S, M, D, L, Lq, P)
    return torch.from_numpy(gv), torch.from_numpy(gl), torch.from_numpy(gw)
